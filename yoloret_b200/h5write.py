"""Minimal pure-Python WRITER for Keras *weights-only* HDF5 files - the counterpart of ``h5lite`` (the reader).

The reference checkpoints with ``model.save_weights(path)`` / ``ModelCheckpoint(save_weights_only=True)``
(reference ``code/train.py:74-79,182-186``), which needs h5py + Keras.  Neither exists in this image, so this module
emits the same on-disk subset ``h5lite`` parses and libhdf5 writes by default for such files (HDF5 File Format
Specification, "version 0" structures): superblock v0, v1 object headers, groups as v1 B-trees ("TREE") of symbol-table
nodes ("SNOD") with names in a local heap ("HEAP"), contiguous little-endian float32 datasets without filters, and v1
attribute messages for Keras' bookkeeping (``layer_names``, ``backend``, ``keras_version`` on the root group,
``weight_names`` on every layer group).  Tree layout of a Keras weights file:

    /                       attrs: layer_names [S], backend, keras_version
    /<layer>                attrs: weight_names [S] = "<layer>/<weight>:0", ...
    /<layer>/<layer>/<weight>:0     dataset

Round trip ``save_keras_weights`` -> ``h5lite.load_keras_weights`` is bit-exact (tests/test_cpu_train.py).  h5py is
not installable here, so reading these files with libhdf5 itself is untested; the structures follow the specification.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _INTERNAL_K = 4, 16          # libhdf5 defaults: <= 8 symbols per SNOD, <= 32 children per TREE node


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _File:
    def __init__(self):
        self.buf = bytearray(96)      # superblock v0, filled in at the end

    def alloc(self, data: bytes) -> int:
        addr = _pad8(len(self.buf))
        self.buf.extend(b"\0" * (addr - len(self.buf)))
        self.buf.extend(data)
        return addr

    def reserve(self, n: int) -> int:
        return self.alloc(b"\0" * n)


# ---- messages -----------------------------------------------------------------------------------------------
def _msg(mtype: int, payload: bytes) -> bytes:
    payload = payload + b"\0" * (_pad8(len(payload)) - len(payload))
    return struct.pack("<HHB3x", mtype, len(payload), 0) + payload


def _dataspace(shape: Sequence[int]) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _dtype_f32() -> bytes:
    # class 1 (floating point), version 1; little-endian, mantissa normalisation 2 (implied leading 1), sign bit 31
    return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _dtype_str(n: int) -> bytes:
    # class 3 (string), version 1; null-padded, ASCII
    return struct.pack("<BBBBI", 0x13, 0x01, 0x00, 0x00, n)


def _attribute(name: str, dtype: bytes, shape: Sequence[int], data: bytes) -> bytes:
    nm = name.encode() + b"\0"
    ds = _dataspace(shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dtype), len(ds))
    body += nm + b"\0" * (_pad8(len(nm)) - len(nm))
    body += dtype + b"\0" * (_pad8(len(dtype)) - len(dtype))
    body += ds + b"\0" * (_pad8(len(ds)) - len(ds))
    body += data
    if len(body) > 0xFFF0:
        raise ValueError("attribute %s is too large for one HDF5 v1 header message (%d bytes)" % (name, len(body)))
    return _msg(0x0C, body)


def _str_array_attr(name: str, values: Sequence[str]) -> bytes:
    enc = [v.encode() for v in values]
    width = max([len(e) for e in enc] + [1])
    return _attribute(name, _dtype_str(width), [len(enc)], b"".join(e.ljust(width, b"\0") for e in enc))


def _str_scalar_attr(name: str, value: str) -> bytes:
    e = value.encode()
    return _attribute(name, _dtype_str(max(len(e), 1)), [], e or b"\0")


def _object_header(f: _File, messages: List[bytes]) -> int:
    body = b"".join(messages)
    return f.alloc(struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body)


# ---- datasets and groups ------------------------------------------------------------------------------------------
def _write_dataset(f: _File, arr: np.ndarray) -> int:
    a = np.ascontiguousarray(arr, dtype="<f4")
    data_addr = f.alloc(a.tobytes()) if a.size else _UNDEF
    msgs = [
        _msg(0x01, _dataspace(a.shape)),
        _msg(0x03, _dtype_f32()),
        _msg(0x05, struct.pack("<BBBB", 2, 2, 0, 0)),                      # fill value v2: late allocation, undefined
        _msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr, a.size * 4)),     # layout v3, contiguous
    ]
    return _object_header(f, msgs)


def _write_group(f: _File, children: List[Tuple[str, int]], attr_msgs: List[bytes]) -> Tuple[int, int, int]:
    """children: (name, object header address).  Returns (object header, B-tree, heap) addresses."""
    children = sorted(children, key=lambda c: c[0].encode())
    # local heap: offset 0 = the empty string, then the names
    seg = bytearray(8)
    offs = []
    for name, _ in children:
        offs.append(len(seg))
        e = name.encode() + b"\0"
        seg.extend(e + b"\0" * (_pad8(len(e)) - len(e)))
    seg_addr = f.alloc(bytes(seg))
    heap_addr = f.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), _UNDEF, seg_addr))
    # symbol-table nodes of <= 2 * leaf K entries each
    per = 2 * _LEAF_K
    nodes = []  # (address, heap offset of the largest name)
    for i in range(0, len(children), per):
        chunk = list(zip(offs[i:i + per], children[i:i + per]))
        body = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
        for off, (_name, ohdr) in chunk:
            body += struct.pack("<QQII16x", off, ohdr, 0, 0)
        body += b"\0" * (8 + per * 40 - len(body))
        nodes.append((f.alloc(body), chunk[-1][0]))
    # B-tree over the nodes (type 0 = group nodes); each level packs <= 2 * internal K children
    level = 0
    fan = 2 * _INTERNAL_K
    if not nodes:
        nodes = []
    while True:
        parents = []
        groups = [nodes[i:i + fan] for i in range(0, len(nodes), fan)] or [[]]
        for gi, grp in enumerate(groups):
            body = b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), _UNDEF, _UNDEF)
            body += struct.pack("<Q", 0)                                    # key 0: the empty string
            for addr, last in grp:
                body += struct.pack("<QQ", addr, last)                      # child i, key i+1 = its largest name
            body += b"\0" * (24 + (2 * fan + 1) * 8 - len(body))
            parents.append((f.alloc(body), grp[-1][1] if grp else 0))
        if len(parents) == 1:
            btree_addr = parents[0][0]
            break
        nodes, level = parents, level + 1
    ohdr = _object_header(f, [_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))] + attr_msgs)
    return ohdr, btree_addr, heap_addr


def save_keras_weights(path: str, weights: Mapping[str, np.ndarray], layer_order: Optional[Sequence[str]] = None,
                       keras_version: str = "2.4.0", backend: str = "tensorflow") -> None:
    """``weights``: ``{"<layer>/<weight>": array}`` (the form ``h5lite.load_keras_weights`` returns).  ``layer_order``:
    layer names in model order (default: first appearance in ``weights``); layers without weights may be listed and get
    an empty group, as Keras writes them."""
    by_layer: Dict[str, List[Tuple[str, np.ndarray]]] = {}
    order: List[str] = list(layer_order) if layer_order is not None else []
    for k, v in weights.items():
        layer, _, leaf = k.partition("/")
        if not leaf or "/" in leaf:
            raise ValueError("weight names must be '<layer>/<weight>', got %r" % k)
        if layer not in by_layer:
            by_layer[layer] = []
            if layer not in order:
                order.append(layer)
        by_layer[layer].append((leaf, np.asarray(v)))
    f = _File()
    top: List[Tuple[str, int]] = []
    for layer in order:
        ws = by_layer.get(layer, [])
        names = ["%s/%s:0" % (layer, leaf) for leaf, _ in ws]
        kids: List[Tuple[str, int]] = []
        if ws:
            inner = [("%s:0" % leaf, _write_dataset(f, arr)) for leaf, arr in ws]
            inner_hdr, _, _ = _write_group(f, inner, [])
            kids.append((layer, inner_hdr))
        attr = _str_array_attr("weight_names", names) if names else \
            _attribute("weight_names", _dtype_str(1), [0], b"")
        hdr, _, _ = _write_group(f, kids, [attr])
        top.append((layer, hdr))
    root_attrs = [_str_array_attr("layer_names", order), _str_scalar_attr("backend", backend),
                  _str_scalar_attr("keras_version", keras_version)]
    root_hdr, root_btree, root_heap = _write_group(f, top, root_attrs)
    eof = _pad8(len(f.buf))
    f.buf.extend(b"\0" * (eof - len(f.buf)))
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)   # root symbol-table entry
    assert len(sb) == 96
    f.buf[:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(f.buf))
