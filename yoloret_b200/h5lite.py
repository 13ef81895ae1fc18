"""Minimal pure-Python reader for Keras *weights-only* HDF5 files.

The reference loads its checkpoints with ``model.load_weights(path)``
(reference ``code/yolo.py:87``, ``code/yolo3/utils.py:389-391``), which needs
h5py + Keras.  Neither exists in this image, so this module parses the small
HDF5 subset those files use (superblock v0, v1 object headers, v1 group
B-trees, contiguous little-endian f32 datasets, no filters) with ``struct`` +
``numpy`` only.  Layout of the subset: SURVEY.md Appendix A.

It is a file-format parser, not arithmetic: both the product path
(``yoloret_b200.weights``) and the CPU oracle read checkpoints through it.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class _Node:
    __slots__ = ("addr", "is_group", "btree", "heap", "shape", "dtype", "data_addr", "nbytes", "attrs")

    def __init__(self, addr):
        self.addr = addr
        self.is_group = False
        self.btree = self.heap = None
        self.shape = None
        self.dtype = None
        self.data_addr = None
        self.nbytes = None
        self.attrs = {}


class H5File:
    """Read-only view of a Keras weights HDF5 file."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != _SIG:
            raise H5Error("not an HDF5 file: %s" % path)
        if b[8] != 0:
            raise H5Error("unsupported superblock version %d" % b[8])
        if b[13] != 8 or b[14] != 8:
            raise H5Error("only 8-byte offsets/lengths supported")
        # superblock v0: root symbol-table entry at byte 56; +8 = object header address
        (root_hdr,) = struct.unpack_from("<Q", b, 56 + 8)
        self.root = self._read_object(root_hdr)

    # ---- object headers -------------------------------------------------
    def _messages(self, addr: int):
        b = self.buf
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error("object header v%d unsupported" % ver)
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                payload = pos + 8
                seen += 1
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", b, payload)
                    blocks.append((caddr, clen))
                else:
                    yield mtype, payload, msize
                pos = payload + msize

    def _read_object(self, addr: int) -> _Node:
        b = self.buf
        node = _Node(addr)
        for mtype, p, msize in self._messages(addr):
            if mtype == 0x11:  # symbol table => group
                node.is_group = True
                node.btree, node.heap = struct.unpack_from("<QQ", b, p)
            elif mtype == 0x01:  # dataspace
                node.shape = self._dataspace(p)
            elif mtype == 0x03:  # datatype
                node.dtype = self._datatype(p)
            elif mtype == 0x08:  # layout
                ver, cls = b[p], b[p + 1]
                if ver != 3:
                    raise H5Error("layout v%d unsupported" % ver)
                if cls == 1:  # contiguous
                    node.data_addr, node.nbytes = struct.unpack_from("<QQ", b, p + 2)
                elif cls == 0:  # compact
                    (sz,) = struct.unpack_from("<H", b, p + 2)
                    node.data_addr, node.nbytes = p + 4, sz
                else:
                    raise H5Error("chunked datasets unsupported")
            elif mtype == 0x0B:
                raise H5Error("filtered datasets unsupported")
            elif mtype == 0x0C:  # attribute
                try:
                    name, val = self._attribute(p)
                    node.attrs[name] = val
                except H5Error:
                    pass
        return node

    def _dataspace(self, p: int) -> Tuple[int, ...]:
        b = self.buf
        ver, rank = b[p], b[p + 1]
        off = p + 8 if ver == 1 else p + 4
        return tuple(struct.unpack_from("<%dQ" % rank, b, off)) if rank else ()

    def _datatype(self, p: int):
        b = self.buf
        cls = b[p] & 0x0F
        (size,) = struct.unpack_from("<I", b, p + 4)
        if cls == 1:
            return np.dtype("<f%d" % size)
        if cls == 0:
            signed = (b[p + 1] >> 3) & 1
            return np.dtype("<%s%d" % ("i" if signed else "u", size))
        if cls == 3:
            return np.dtype("S%d" % size)
        raise H5Error("datatype class %d unsupported" % cls)

    def _attribute(self, p: int):
        b = self.buf
        ver = b[p]
        if ver != 1:
            raise H5Error("attribute v%d unsupported" % ver)
        nsz, tsz, ssz = struct.unpack_from("<HHH", b, p + 2)
        pad = lambda n: (n + 7) & ~7
        q = p + 8
        name = b[q:q + nsz].split(b"\0")[0].decode()
        q += pad(nsz)
        dt = self._datatype(q)
        q += pad(tsz)
        shape = self._dataspace(q)
        q += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(b, dtype=dt, count=n, offset=q).reshape(shape)
        if dt.kind == "S":
            arr = np.array([s.split(b"\0")[0].decode() for s in arr.ravel()], dtype=object).reshape(shape)
        return name, arr

    # ---- groups ----------------------------------------------------------
    def _heap_name(self, heap_addr: int, off: int) -> str:
        b = self.buf
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5Error("bad local heap")
        (seg,) = struct.unpack_from("<Q", b, heap_addr + 24)
        end = b.index(b"\0", seg + off)
        return b[seg + off:end].decode()

    def _walk_btree(self, addr: int, heap: int, out: List[Tuple[str, int]]):
        b = self.buf
        if addr == _UNDEF:
            return
        sig = b[addr:addr + 4]
        if sig == b"TREE":
            _ntype, _level, nent = struct.unpack_from("<BBH", b, addr + 4)
            for i in range(nent):
                (child,) = struct.unpack_from("<Q", b, addr + 24 + 16 * i + 8)
                self._walk_btree(child, heap, out)
        elif sig == b"SNOD":
            (cnt,) = struct.unpack_from("<H", b, addr + 6)
            for i in range(cnt):
                noff, ohdr = struct.unpack_from("<QQ", b, addr + 8 + 40 * i)
                out.append((self._heap_name(heap, noff), ohdr))
        else:
            raise H5Error("bad B-tree node signature %r" % sig)

    def children(self, node: _Node) -> Dict[str, _Node]:
        if not node.is_group:
            return {}
        ents: List[Tuple[str, int]] = []
        self._walk_btree(node.btree, node.heap, ents)
        return {name: self._read_object(addr) for name, addr in ents}

    def read(self, node: _Node) -> np.ndarray:
        if node.data_addr is None or node.dtype is None:
            raise H5Error("not a dataset")
        n = int(np.prod(node.shape)) if node.shape else 1
        if node.data_addr == _UNDEF:
            return np.zeros(node.shape, node.dtype)
        return np.frombuffer(self.buf, dtype=node.dtype, count=n, offset=node.data_addr).reshape(node.shape).copy()

    # ---- Keras view --------------------------------------------------------
    def layer_names(self) -> List[str]:
        names = self.root.attrs.get("layer_names")
        return [str(s) for s in names.ravel()] if names is not None else []

    def weights(self) -> Dict[str, np.ndarray]:
        """Flat ``{"<layer>/<weight>": array}`` map, e.g. ``"Conv1/kernel"``."""
        out: Dict[str, np.ndarray] = {}

        def rec(node: _Node, prefix: str):
            for name, ch in self.children(node).items():
                if ch.is_group:
                    rec(ch, prefix + [name] if False else prefix + "/" + name if prefix else name)
                else:
                    out[(prefix + "/" + name) if prefix else name] = self.read(ch)

        rec(self.root, "")
        flat: Dict[str, np.ndarray] = {}
        for k, v in out.items():
            parts = k.split("/")
            # /<layer>/<layer>/<weight>:0  (nested models add more levels; keep layer + leaf)
            layer, leaf = parts[0], parts[-1]
            if leaf.endswith(":0"):
                leaf = leaf[:-2]
            flat[layer + "/" + leaf] = v
        return flat


def load_keras_weights(path: str) -> Dict[str, np.ndarray]:
    return H5File(path).weights()
