"""CPU: the built-in HDF5 subset (efficientdet_b200/utils/hdf5.py) behind load_weights / save_weights of Keras
`.h5` weight files (train.py:329-332, utils/train.py:10-35).  The READER is pinned on the one file written by
libhdf5 that exists in this image (a MATLAB 7.4 v7.3 MAT-file = HDF5 behind a 512-byte user block, from scipy's
test data, BSD licence, copied to tests/golden/): user block + superblock v0, symbol-table group (v1 B-tree,
SNOD, local heap), v1 object header, dataspace, IEEE float datatype, layout message v1/2, fixed-string attribute.
No h5py-written Keras file is available, so the Keras LAYOUT (layer_names / weight_names conventions) and the
writer are checked by round trips + byte-level known answers derived from the format specification."""
import struct

import numpy as np
import pytest

from efficientdet_b200.utils import hdf5


def _weights(n_layers=70, seed=0):
    rng = np.random.default_rng(seed)
    w = {}
    for i in range(n_layers):
        L = "block%da_conv_%d" % (i % 7 + 1, i)
        w[L + "/kernel"] = rng.standard_normal((3, 3, 4, 8)).astype(np.float32)
        w[L + "/bias"] = rng.standard_normal(8).astype(np.float32)
    w["stem_bn/gamma"] = np.ones(32, np.float32)
    w["stem_bn/moving_variance"] = rng.uniform(0.5, 2, 32).astype(np.float32)
    w["box_head/regress_head_conv_0/kernel"] = rng.standard_normal((3, 3, 8, 8)).astype(np.float32)
    w["w_bi_fpn_add_3/w_bi_fpn_add_3"] = np.array([0.5, 0.25, 0.25], np.float32)
    w["boxes/anchor_boxes_baked"] = rng.uniform(0, 512, (1, 49104, 4)).astype(np.float32)
    w["scalar_layer/step"] = np.array(7, np.int64)
    return w


def test_reader_on_a_file_written_by_libhdf5():
    """scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat: MATLAB 7.4 saved `testdouble = 0:pi/4:2*pi` with
    libhdf5 (scipy's own tests list this variable's value in test_mio.py)."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    g = hdf5.open_file(os.path.join(here, "golden", "libhdf5_matlab74_testdouble.mat"))
    assert g.keys() == ["testdouble"]
    v = g["testdouble"]
    assert v.shape == (9, 1) and v.dtype == np.float64
    assert np.array_equal(v.ravel(), np.pi / 4 * np.arange(9))
    assert g._r.attributes(g._links["testdouble"]) == {"MATLAB_class": b"double"}
    assert g._r.base == 512 and (g._r.leaf_k, g._r.internal_k) == (4, 16)
    # the writer encodes the structures both files share byte for byte like libhdf5 did in that file
    msgs = {t: d for t, _f, d in g._r.messages(g._links["testdouble"])}
    assert msgs[0x0003][:20] == hdf5._Writer._datatype(np.float64)
    assert msgs[0x0001] == hdf5._Writer._dataspace((9, 1))
    mine = hdf5._Writer()._attribute("MATLAB_class", np.bytes_(b"double"))[8:]      # message body
    theirs = bytearray(msgs[0x000C])
    assert theirs[25] == 0x00 and mine[25] == 0x01        # string padding: libhdf5/MATLAB null-terminated, numpy 'S' null-padded
    theirs[25] = 0x01
    assert bytes(theirs) == mine


def test_keras_weight_file_round_trip(tmp_path):
    w = _weights()
    p = str(tmp_path / "w.h5")
    hdf5.save_keras_weights(p, w)
    got = hdf5.load_keras_weights(p)
    assert set(got) == set(w)
    for k in w:
        assert got[k].dtype == w[k].dtype and got[k].shape == w[k].shape, k
        assert np.array_equal(got[k], w[k]), k
    root = hdf5.open_file(p)
    at = root.attrs
    assert at["backend"] == b"tensorflow" and at["keras_version"] == b"2.2.4-tf"
    layers = [x.decode() for x in at["layer_names"]]
    assert len(layers) == 70 + 5 and sorted(root.keys()) == sorted(layers)     # 10 symbol-table nodes under one B-tree node
    names = [x.decode() for x in root["box_head"].attrs["weight_names"]]
    assert names == ["box_head/regress_head_conv_0/kernel:0"]
    assert np.array_equal(root["box_head/box_head/regress_head_conv_0/kernel:0"], w["box_head/regress_head_conv_0/kernel"])


def test_nested_model_weight_names_without_outer_prefix(tmp_path):
    """tf.keras names the weights of a nested Model after the inner layer; load prefixes the outer layer."""
    wr = hdf5._Writer()
    k = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    inner = wr.group({"kernel:0": wr.dataset(k)})[0]
    head = wr.group({"regress_head_conv_0": inner}, attrs={"weight_names": ["regress_head_conv_0/kernel:0"]})[0]
    data = wr.finish(wr.group({"box_head": head}, attrs={"layer_names": ["box_head"], "backend": "tensorflow"}))
    p = tmp_path / "n.h5"
    p.write_bytes(data)
    got = hdf5.load_keras_weights(str(p))
    assert list(got) == ["box_head/regress_head_conv_0/kernel"] and np.array_equal(got[list(got)[0]], k)


def test_many_children_use_a_multi_level_btree(tmp_path):
    wr = hdf5._Writer()
    ch = {"layer_%04d" % i: wr.dataset(np.full((2,), i, np.int32)) for i in range(700)}   # > 2*4*2*16 = 256 entries
    data = wr.finish(wr.group(ch))
    p = tmp_path / "m.h5"
    p.write_bytes(data)
    root = hdf5.open_file(str(p))
    assert root.keys() == sorted(ch)
    for i in (0, 255, 256, 699):
        assert root["layer_%04d" % i].tolist() == [i, i]


def test_structures_follow_the_format_specification(tmp_path):
    p = str(tmp_path / "s.h5")
    hdf5.save_keras_weights(p, {"a/kernel": np.array([[1.5, -2.0]], np.float32)})
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0                 # signature, superblock version 0
    assert b[13] == 8 and b[14] == 8                                      # sizes of offsets / lengths
    leaf_k, internal_k = struct.unpack_from("<HH", b, 16)
    assert (leaf_k, internal_k) == (4, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and eof == len(b) and free == drv == 0xFFFFFFFFFFFFFFFF
    name_off, root, cache = struct.unpack_from("<QQI", b, 56)
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert cache == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    assert b[root] == 1                                                   # version-1 object header
    assert len(b) % 8 == 0 and root % 8 == 0 and btree % 8 == 0
    # the B-tree node has the full size the superblock's K implies: 24 + (2K+1)*8 + 2K*8
    snod, = struct.unpack_from("<Q", b, btree + 24 + 8)
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 1
    # float32 datatype message body: class 1 v1, little endian, implied-msb normalisation, sign bit 31,
    # size 4; offset 0, precision 32, exponent at 23 (8 bits), mantissa at 0 (23 bits), bias 127
    assert hdf5._Writer._datatype(np.float32) == bytes([0x11, 0x20, 31, 0, 4, 0, 0, 0, 0, 0, 32, 0, 23, 8, 0, 23,
                                                        127, 0, 0, 0])
    assert hdf5._Writer._dataspace((3, 5)) == bytes([1, 2, 0, 0, 0, 0, 0, 0]) + struct.pack("<QQ", 3, 5)


def test_reader_handles_attribute_versions_vlen_strings_compact_and_chunked(tmp_path):
    """Structures h5py can emit that the writer does not: version-3 attributes, variable-length strings in the
    global heap (str attributes of tf.keras >= 2.3), compact and unfiltered chunked layouts, continuation blocks.
    Built by hand from the specification."""
    wr = hdf5._Writer()
    # global heap with one object "tensorflow"
    s = b"tensorflow"
    obj = struct.pack("<HHIQ", 1, 1, 0, len(s)) + s + b"\0" * (-len(s) % 8)
    gcol_size = 16 + len(obj) + 16
    gcol = wr._append(b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, gcol_size) + obj + struct.pack("<HHIQ", 0, 0, 0, 0))
    vlen_dt = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x13, 0x00, 0, 0, 1)
    scalar_ds = struct.pack("<BBBB", 2, 0, 0, 0)                        # dataspace v2, rank 0, type scalar
    nm = b"backend\0"
    attr3 = struct.pack("<BBHHHB", 3, 0, len(nm), len(vlen_dt), len(scalar_ds), 1) + nm + vlen_dt + scalar_ds + \
        struct.pack("<IQI", len(s), gcol, 1)
    # compact dataset
    vals = np.array([3, 1, 4, 1, 5], np.int16)
    compact = wr._object_header([wr._msg(1, wr._dataspace(vals.shape)), wr._msg(3, wr._datatype(vals.dtype)),
                                 wr._msg(8, struct.pack("<BBH", 3, 0, vals.nbytes) + vals.tobytes())])
    # chunked 4x6 float32 dataset in 2x4 chunks (edge chunks padded), no filters
    a = np.arange(24, dtype=np.float32).reshape(4, 6)
    keys = []
    for oy in (0, 2):
        for ox in (0, 4):
            c = np.zeros((2, 4), np.float32)
            blk = a[oy:oy + 2, ox:ox + 4]
            c[:blk.shape[0], :blk.shape[1]] = blk
            keys.append(((oy, ox), wr._append(c.tobytes())))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), hdf5.UNDEF, hdf5.UNDEF)
    for (oy, ox), addr in keys:
        node += struct.pack("<IIQQQ", 32, 0, oy, ox, 0) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQQ", 0, 0, 4, 6, 0)
    tree = wr._append(node)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", tree) + struct.pack("<III", 2, 4, 4)
    # the chunked dataset's header uses a continuation block for its layout message
    cont_body = wr._msg(8, layout)
    cont = wr._append(cont_body)
    chunked = wr._object_header([wr._msg(1, wr._dataspace(a.shape)), wr._msg(3, wr._datatype(a.dtype)),
                                 wr._msg(0x10, struct.pack("<QQ", cont, len(cont_body)))])
    # patch the message count of that header (2 + continuation + 1 message inside the continuation block)
    struct.pack_into("<H", wr.buf, chunked + 2, 4)
    root = wr.group({"c": compact, "k": chunked})
    # splice the hand-made attribute into the root header through a continuation block as well
    rb = wr._msg(0x000C, attr3)
    ra = wr._append(rb)
    hdr = root[0]
    nmsg, = struct.unpack_from("<H", wr.buf, hdr + 2)
    # reuse: build a new root header = symbol table message + continuation
    new_root = wr._object_header([wr._msg(0x0011, struct.pack("<QQ", root[1], root[2])),
                                  wr._msg(0x10, struct.pack("<QQ", ra, len(rb)))])
    struct.pack_into("<H", wr.buf, new_root + 2, 3)
    p = tmp_path / "h.h5"
    p.write_bytes(wr.finish((new_root, root[1], root[2])))
    g = hdf5.open_file(str(p))
    assert g.attrs == {"backend": "tensorflow"}
    assert g["c"].tolist() == [3, 1, 4, 1, 5] and g["c"].dtype == np.int16
    assert np.array_equal(g["k"], a)


def test_unsupported_files_fail_loudly(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(ValueError):
        hdf5.open_file(str(p))
    p.write_bytes(hdf5.SIGNATURE + bytes([2]) + b"\0" * 64)
    with pytest.raises(NotImplementedError):
        hdf5.open_file(str(p))
    # truncated file: structures point past the end -> an exception, never a hang or garbage
    good = tmp_path / "g.h5"
    hdf5.save_keras_weights(str(good), {"a/kernel": np.ones((4, 4), np.float32)})
    data = good.read_bytes()
    p.write_bytes(data[:len(data) // 2])
    with pytest.raises((ValueError, struct.error, IndexError, KeyError)):
        hdf5.load_keras_weights(str(p))
