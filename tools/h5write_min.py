"""Minimal HDF5 WRITER for test input files (no h5py / HDF5 library exists in this image).

Writes the classic on-disk format the HDF5 library produces with its default settings -- superblock version 0,
version-1 object headers, old-style groups (symbol-table message, one version-1 B-tree leaf node and one symbol
node per group, names in a local heap), dataspace v1, datatype v1, data layout v3 -- for the subset parPE's
input files need: nested groups; float64 / int32 / int64 datasets, contiguous or chunked (optionally
gzip-compressed, as ``python/parpe/hdf5_pe_input.py`` writes them); scalar numeric attributes.

``write_parpe_input(path, tables, hier=None)`` lays a ``ProblemTables`` out as a parPE parameter-estimation input
file (doc/hdf5.md:1-172).  Used by tests/test_h5_reader.py; the reader under test (parpe_b200/csrc/h5min.h) is
separately pinned on a file written by the real HDF5 library.

Group B-trees: the group leaf-node K in the superblock is set large enough for every group's entries to sit in ONE
symbol node (2K entries); that is a legal choice of the file-creation parameters.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * ((-len(b)) % 8)


class Node:
    def __init__(self):
        self.children = {}   # name -> Node | Dataset
        self.attrs = {}      # name -> numpy scalar/array


class Dataset:
    def __init__(self, data, chunks=None, compress=False):
        self.data = np.ascontiguousarray(data)
        self.chunks = chunks
        self.compress = compress
        self.attrs = {}


class VlenDataset:
    """1-D dataset of variable-length int32 sequences (h5py.vlen_dtype(np.int32)), e.g. parPE's
    /fixedParameters/reinitializationIndices (python/parpe/hdf5_pe_input.py:430-437) -- or, with ``strings``, of
    variable-length strings (what h5py writes for a list of str, e.g. /parameters/parameterNames)."""

    def __init__(self, sequences, strings=False):
        self.strings = strings
        if strings:
            self.sequences = [np.frombuffer(s.encode(), dtype=np.uint8) for s in sequences]
        else:
            self.sequences = [np.ascontiguousarray(s, dtype=np.int32) for s in sequences]
        self.attrs = {}


class H5Writer:
    def __init__(self):
        self.root = Node()
        self.buf = bytearray()

    # ---- tree building -----------------------------------------------------------------
    def group(self, path) -> Node:
        node = self.root
        for part in [p for p in path.split("/") if p]:
            node = node.children.setdefault(part, Node())
        return node

    def dataset(self, path, data, chunks=None, compress=False):
        parts = [p for p in path.split("/") if p]
        g = self.group("/".join(parts[:-1]))
        g.children[parts[-1]] = Dataset(data, chunks, compress)

    def vlen_dataset(self, path, sequences, strings=False):
        parts = [p for p in path.split("/") if p]
        g = self.group("/".join(parts[:-1]))
        g.children[parts[-1]] = VlenDataset(sequences, strings)

    def attr(self, path, name, value):
        self.group(path).attrs[name] = np.asarray(value)

    # ---- low-level emitters ----------------------------------------------------------------
    def _alloc(self, data: bytes, align=8) -> int:
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _datatype(dt: np.dtype) -> bytes:
        if dt.kind == "f":
            # class 1 (float) version 1; bit field: little endian, mantissa normalisation "implied", sign bit position
            size = dt.itemsize
            if size == 8:
                bits = bytes([0x20, 63, 0])
                props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            else:
                bits = bytes([0x20, 31, 0])
                props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            return bytes([0x11]) + bits + struct.pack("<I", size) + props
        if dt.kind in "iu":
            bits = bytes([0x08 if dt.kind == "i" else 0x00, 0, 0])
            return bytes([0x10]) + bits + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
        raise TypeError(dt)

    @staticmethod
    def _dataspace(shape) -> bytes:
        return bytes([1, len(shape), 0, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", int(n)) for n in shape)

    @staticmethod
    def _message(mtype: int, body: bytes) -> bytes:
        body = _pad8(body)
        return struct.pack("<HHBBBB", mtype, len(body), 0, 0, 0, 0) + body

    def _attr_message(self, name: str, value: np.ndarray) -> bytes:
        nm = name.encode() + b"\0"
        dt = self._datatype(value.dtype)
        ds = self._dataspace(value.shape)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + value.tobytes()
        return self._message(0x000C, body)

    def _object_header(self, messages: list) -> int:
        body = b"".join(messages)
        hdr = struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4
        return self._alloc(hdr + body)

    def _write_vlen_dataset(self, d: VlenDataset) -> int:
        """Elements (length u32, global heap collection address u64, object index u32); the sequences are the
        objects 1..n of ONE global heap collection (GCOL version 1), object 0 = the free space at its end."""
        objs = bytearray()
        for i, s in enumerate(d.sequences):
            if len(s):
                objs += struct.pack("<HHIQ", i + 1, 1, 0, s.nbytes) + _pad8(s.tobytes())
        total = max(4096, 16 + len(objs) + 16)
        total += (-total) % 8
        free = total - 16 - len(objs)
        coll = b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", total) + bytes(objs)
        coll += struct.pack("<HHIQ", 0, 0, 0, free) + b"\0" * (free - 16)
        assert len(coll) == total
        caddr = self._alloc(coll)
        elems = b"".join(struct.pack("<IQI", len(s), caddr if len(s) else 0, i + 1 if len(s) else 0)
                         for i, s in enumerate(d.sequences))
        if d.strings:
            # class 9, type "string" (bit field: type 1, padding null-terminated, ASCII); base type: 1-byte string
            base = bytes([0x13, 0x00, 0x00, 0x00]) + struct.pack("<I", 1)
            dtype = bytes([0x19, 0x01, 0x00, 0x00]) + struct.pack("<I", 16) + base
        else:
            base = self._datatype(np.dtype(np.int32))
            dtype = bytes([0x19, 0x00, 0x00, 0x00]) + struct.pack("<I", 16) + base
        msgs = [self._message(0x0001, self._dataspace((len(d.sequences),))), self._message(0x0003, dtype)]
        addr = self._alloc(elems) if elems else UNDEF
        msgs.append(self._message(0x0008, struct.pack("<BBQQ", 3, 1, addr, len(elems))))
        return self._object_header(msgs)

    def _write_dataset(self, d: Dataset) -> int:
        a = d.data
        shape = a.shape if a.ndim else (1,)
        msgs = [self._message(0x0001, self._dataspace(shape)), self._message(0x0003, self._datatype(a.dtype))]
        es = a.dtype.itemsize
        if d.chunks is None:
            addr = self._alloc(a.tobytes()) if a.size else UNDEF
            msgs.append(self._message(0x0008, struct.pack("<BBQQ", 3, 1, addr, a.size * es)))
        else:
            chunks = tuple(int(c) for c in d.chunks)
            assert len(chunks) == len(shape)
            a2 = a.reshape(shape)
            # every chunk, padded to the full chunk shape, (optionally) deflated
            keys = []
            grid = [range(0, shape[k], chunks[k]) for k in range(len(shape))]
            import itertools
            for off in itertools.product(*grid):
                block = np.zeros(chunks, dtype=a.dtype)
                sl = tuple(slice(off[k], min(off[k] + chunks[k], shape[k])) for k in range(len(shape)))
                part = a2[sl]
                block[tuple(slice(0, s) for s in part.shape)] = part
                raw = block.tobytes()
                if d.compress:
                    raw = zlib.compress(raw, 4)
                caddr = self._alloc(raw)
                keys.append((len(raw), off, caddr))
            # one leaf node of the chunk B-tree (type 1)
            rank = len(shape)
            node = b"TREE" + bytes([1, 0]) + struct.pack("<H", len(keys)) + struct.pack("<QQ", UNDEF, UNDEF)
            for (csize, off, caddr) in keys:
                node += struct.pack("<II", csize, 0) + b"".join(struct.pack("<Q", o) for o in off) + struct.pack("<Q", 0)
                node += struct.pack("<Q", caddr)
            last = tuple(shape[k] + 0 for k in range(rank))
            node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", o) for o in last) + struct.pack("<Q", 0)
            baddr = self._alloc(node)
            body = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", baddr)
            body += b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", es)
            msgs.append(self._message(0x0008, body))
            if d.compress:
                # filter pipeline v1: one filter, id 1 (deflate), name "deflate", one client value (level)
                name = _pad8(b"deflate\0")
                fp = struct.pack("<BB", 1, 1) + b"\0" * 6 + struct.pack("<HHHH", 1, len(name), 1, 1) + name + struct.pack("<II", 4, 0)
                msgs.append(self._message(0x000B, fp))
        for k, v in d.attrs.items():
            msgs.append(self._attr_message(k, v))
        return self._object_header(msgs)

    def _write_group(self, g: Node) -> int:
        entries = []
        for name in sorted(g.children):
            ch = g.children[name]
            if isinstance(ch, VlenDataset):
                addr = self._write_vlen_dataset(ch)
            else:
                addr = self._write_dataset(ch) if isinstance(ch, Dataset) else self._write_group(ch)[0]
            entries.append((name, addr))
        # local heap: offset 0 holds the empty string (the B-tree's first key), then the names
        heap = bytearray(b"\0" * 8)
        offs = []
        for name, _ in entries:
            offs.append(len(heap))
            heap += _pad8(name.encode() + b"\0")
        heap_data = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap), UNDEF, heap_data))
        # symbol node with all entries, B-tree leaf with this one child
        snod = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(entries))
        for (name, addr), off in zip(entries, offs):
            snod += struct.pack("<QQII", off, addr, 0, 0) + b"\0" * 16
        snod_addr = self._alloc(snod)
        if entries:
            tree = b"TREE" + bytes([0, 0]) + struct.pack("<H", 1) + struct.pack("<QQ", UNDEF, UNDEF)
            tree += struct.pack("<Q", 0) + struct.pack("<Q", snod_addr) + struct.pack("<Q", offs[-1])
        else:
            tree = b"TREE" + bytes([0, 0]) + struct.pack("<H", 0) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<Q", 0)
        tree_addr = self._alloc(tree)
        msgs = [self._message(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        for k, v in g.attrs.items():
            msgs.append(self._attr_message(k, v))
        return self._object_header(msgs), tree_addr, heap_addr

    def _count_max_children(self, g: Node) -> int:
        n = len(g.children)
        for ch in g.children.values():
            if isinstance(ch, Node):
                n = max(n, self._count_max_children(ch))
        return n

    def save(self, path):
        self.buf = bytearray(b"\0" * 96)  # superblock (56 bytes + 40-byte root symbol table entry)
        res = self._write_group(self.root)
        root_hdr, tree_addr, heap_addr = res
        leaf_k = max(4, (self._count_max_children(self.root) + 1) // 2 + 1)
        assert leaf_k < 65536
        sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", leaf_k, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
        assert len(sb) == 96
        self.buf[0:96] = sb
        with open(path, "wb") as fh:
            fh.write(bytes(self.buf))


def write_parpe_input(path, tab, hier=None, compress=True, attrs=None):
    """ProblemTables -> parPE parameter-estimation input file (doc/hdf5.md).  Conditions use one column of
    /fixedParameters/k each; pre-equilibration conditions (tab.has_preeq / k_preeq / reinit_idx) become further
    columns of k referenced by simulationConditions column 0, with /fixedParameters/reinitializationIndices as parPE
    writes it (a ragged int32 dataset)."""
    C, np_, nk, ny = tab.n_conditions, tab.np_, tab.nk, tab.ny
    w = H5Writer()
    w.attr("/amiciOptions", "atol", np.float64(tab.atol))
    w.attr("/amiciOptions", "rtol", np.float64(tab.rtol))
    w.attr("/amiciOptions", "maxsteps", np.int32(tab.maxsteps))
    w.attr("/amiciOptions", "stldet", np.int32(tab.stldet))
    w.attr("/amiciOptions", "sensi", np.int32(1))
    w.attr("/amiciOptions", "sensi_meth", np.int32(1))
    for k_, v_ in (attrs or {}).items():
        w.attr("/amiciOptions", k_, v_)
    kmat = np.asarray(tab.k, dtype=np.float64)
    pre_col = np.full(C, -1)
    if getattr(tab, "has_preeq", None) is not None:
        extra = []
        for c in range(C):
            if tab.has_preeq[c]:
                pre_col[c] = C + len(extra)
                extra.append(np.asarray(tab.k_preeq[c], dtype=np.float64))
        if extra:
            kmat = np.vstack([kmat] + extra)
        lists = tab.reinit_idx if tab.reinit_idx is not None else [[] for _ in range(C)]
        w.vlen_dataset("/fixedParameters/reinitializationIndices", [lists[c] if tab.has_preeq[c] else [] for c in range(C)])
    w.dataset("/fixedParameters/k", kmat.T.copy(), chunks=(max(nk, 1), 1), compress=compress)
    sc = np.stack([pre_col, np.arange(C)], axis=1).astype(np.int32)
    w.dataset("/fixedParameters/simulationConditions", sc)
    for c in range(C):
        lo, hi = int(tab.t_offsets[c]), int(tab.t_offsets[c + 1])
        nt = hi - lo
        ch1 = (max(nt, 1),)
        ch2 = (max(nt, 1), ny)
        w.dataset("/measurements/t/%d" % c, tab.t[lo:hi].astype(np.float64), chunks=ch1 if compress else None, compress=compress)
        w.dataset("/measurements/y/%d" % c, tab.y_of(c).astype(np.float64), chunks=ch2 if compress else None, compress=compress)
        w.dataset("/measurements/ysigma/%d" % c, tab.sigma_of(c).astype(np.float64), chunks=ch2 if compress else None,
                  compress=compress)
    w.dataset("/parameters/optimizationSimulationMapping", np.asarray(tab.mapping, dtype=np.int32).T.copy(),
              chunks=(np_, 1), compress=compress)
    w.dataset("/parameters/pscaleSimulation", np.asarray(tab.pscale_sim, dtype=np.int32))
    w.dataset("/parameters/pscaleOptimization", np.asarray(tab.pscale_opt, dtype=np.int32))
    w.dataset("/parameters/lowerBound", np.asarray(tab.lower_bound, dtype=np.float64)
              if getattr(tab, "lower_bound", None) is not None else np.full(tab.n_opt, -5.0))
    w.dataset("/parameters/upperBound", np.asarray(tab.upper_bound, dtype=np.float64)
              if getattr(tab, "upper_bound", None) is not None else np.full(tab.n_opt, 5.0))
    if getattr(tab, "nominal_values", None) is not None:
        w.dataset("/parameters/nominalValues", np.asarray(tab.nominal_values, dtype=np.float64))
    if getattr(tab, "parameter_names", None) is not None:
        w.vlen_dataset("/parameters/parameterNames", list(tab.parameter_names), strings=True)
    if tab.overrides is not None:
        w.dataset("/parameters/parameterOverrides", np.asarray(tab.overrides, dtype=np.float64).T.copy(), chunks=(np_, 1),
                  compress=compress)
    h = hier if hier is not None else tab.hier
    if h is not None:
        for kind in ("scaling", "offset", "sigma"):
            idx = getattr(h, kind + "_idx")
            mp = getattr(h, kind + "_map")
            if len(idx):
                w.dataset("/%sParameterIndices" % kind, np.asarray(idx, dtype=np.int32))
                w.dataset("/%sParametersMapToObservables" % kind, np.asarray(mp, dtype=np.int32).reshape(-1, 3))
    w.save(path)
