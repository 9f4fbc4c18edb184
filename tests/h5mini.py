"""Test helper: a small independent reader for the subset of HDF5 that the reference's Output_writer produces and host/output_h5.h
writes -- superblock version 0, version-1 object headers, group B-tree + symbol-table nodes + local heap, contiguous
little-endian float64 datasets, version-1 attributes.  Follows the published HDF5 File Format Specification; written
separately from the writer (different language, parses by walking the on-disk links) because neither libhdf5 nor h5py exists
in this image."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class H5File:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("bad signature")
        if b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise H5Error("only superblock version 0 with 8-byte offsets / lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        base, _free, eof, _driver = struct.unpack_from("<QQQQ", b, 24)
        if base != 0 or eof != len(b):
            raise H5Error(f"base address {base}, end-of-file address {eof}, file length {len(b)}")
        _name_off, root_header, cache_type = struct.unpack_from("<QQI", b, 56)
        btree, heap = struct.unpack_from("<QQ", b, 80)
        msgs = self._object_header(root_header)
        stab = [m for m in msgs if m[0] == 0x0011]
        if len(stab) != 1:
            raise H5Error("root group has no symbol table message")
        btree2, heap2 = struct.unpack_from("<QQ", stab[0][1], 0)
        if cache_type == 1 and (btree, heap) != (btree2, heap2):
            raise H5Error("root symbol table entry cache disagrees with the object header")
        self.links = {}
        self._walk_btree(btree2, self._heap(heap2))
        self.datasets = {name: self._dataset(addr) for name, addr in self.links.items()}

    # -- structures
    def _heap(self, addr):
        b = self.buf
        if b[addr:addr + 4] != b"HEAP" or b[addr + 4] != 0:
            raise H5Error("bad local heap")
        size, free_head, data = struct.unpack_from("<QQQ", b, addr + 8)
        if data + size > len(b):
            raise H5Error("heap data segment beyond the end of the file")
        while free_head != 1:  # walk the free list: every block inside the segment, at least 16 bytes
            if free_head + 16 > size:
                raise H5Error("bad heap free list")
            nxt, blk = struct.unpack_from("<QQ", b, data + free_head)
            if blk < 16 or free_head + blk > size:
                raise H5Error("bad heap free block")
            free_head = nxt
        return b[data:data + size]

    @staticmethod
    def _name(heap, off):
        end = heap.index(b"\0", off)
        return heap[off:end].decode()

    def _walk_btree(self, addr, heap):
        b = self.buf
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 0:
            raise H5Error("bad group B-tree node")
        level, used = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        if addr + 24 + 16 * 2 * self.internal_k + 8 > len(b):
            raise H5Error("B-tree node beyond the end of the file")
        pos = addr + 24
        prev_key = None
        for i in range(used):
            key_lo, child, key_hi = struct.unpack_from("<QQQ", b, pos)
            pos += 16
            if level > 0:
                self._walk_btree(child, heap)
            else:
                names = self._snod(child, heap)
                # keys bracket the names of the child: key_lo < every name <= key_hi
                if names and not (self._name(heap, key_lo) < names[0] and names[-1] <= self._name(heap, key_hi)):
                    raise H5Error("B-tree keys do not bracket the symbol table node")
            prev_key = key_hi
        return prev_key

    def _snod(self, addr, heap):
        b = self.buf
        if b[addr:addr + 4] != b"SNOD" or b[addr + 4] != 1:
            raise H5Error("bad symbol table node")
        count = struct.unpack_from("<H", b, addr + 6)[0]
        if count > 2 * self.leaf_k or addr + 8 + 40 * 2 * self.leaf_k > len(b):
            raise H5Error("symbol table node overflows")
        names = []
        for i in range(count):
            name_off, header = struct.unpack_from("<QQ", b, addr + 8 + 40 * i)
            name = self._name(heap, name_off)
            names.append(name)
            self.links[name] = header
        if names != sorted(names):
            raise H5Error("symbol table entries are not sorted by name")
        return names

    def _object_header(self, addr):
        b = self.buf
        if b[addr] != 1:
            raise H5Error("only version-1 object headers")
        nmsgs, _refs, size = struct.unpack_from("<HII", b, addr + 2)
        pos, end = addr + 16, addr + 16 + size
        msgs = []
        while pos < end and len(msgs) < nmsgs:
            mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
            if msize % 8:
                raise H5Error("message size not a multiple of 8")
            msgs.append((mtype, b[pos + 8:pos + 8 + msize]))
            pos += 8 + msize
        if pos != end or len(msgs) != nmsgs:
            raise H5Error("object header size / message count mismatch")
        return msgs

    @staticmethod
    def _dataspace(body):
        if body[0] != 1:
            raise H5Error("only version-1 dataspaces")
        rank, flags = body[1], body[2]
        if flags:
            raise H5Error("max dims / permutation not supported")
        return tuple(struct.unpack_from("<%dQ" % rank, body, 8))

    @staticmethod
    def _datatype(body):
        cls, version = body[0] & 0x0F, body[0] >> 4
        size = struct.unpack_from("<I", body, 4)[0]
        if version != 1:
            raise H5Error("only version-1 datatypes")
        if cls == 1:  # floating point
            bits = body[1] | (body[2] << 8) | (body[3] << 16)
            off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", body, 8)
            if (bits & 1) == 0 and size == 8 and (off, prec, eloc, esize, mloc, msize, bias) == (0, 64, 52, 11, 0, 52, 1023) \
                    and ((bits >> 8) & 0xFF) == 63 and ((bits >> 4) & 3) == 2:
                return np.dtype("<f8"), 20
            raise H5Error("floating-point type is not IEEE little-endian float64")
        if cls == 0:  # fixed point
            off, prec = struct.unpack_from("<HH", body, 8)
            if (body[1] & 1) == 0 and (body[1] & 8) and size == 4 and (off, prec) == (0, 32):
                return np.dtype("<i4"), 12
            raise H5Error("fixed-point type is not a little-endian signed int32")
        raise H5Error(f"datatype class {cls} not supported")

    def _attribute(self, body):
        if body[0] != 1:
            raise H5Error("only version-1 attributes")
        name_size, type_size, space_size = struct.unpack_from("<HHH", body, 2)
        pad = lambda n: (n + 7) // 8 * 8  # noqa: E731
        pos = 8
        name = body[pos:pos + name_size - 1].decode()
        pos += pad(name_size)
        dtype, used = self._datatype(body[pos:pos + type_size])
        if used != type_size:
            raise H5Error("attribute datatype size")
        pos += pad(type_size)
        dims = self._dataspace(body[pos:pos + space_size])
        if space_size != 8 + 8 * len(dims):
            raise H5Error("attribute dataspace size")
        pos += pad(space_size)
        count = int(np.prod(dims)) if dims else 1
        value = np.frombuffer(body, dtype=dtype, count=count, offset=pos).copy()
        return name, (value.reshape(dims) if dims else value[0])

    def _dataset(self, addr):
        dims = dtype = layout = None
        attrs = {}
        for mtype, body in self._object_header(addr):
            if mtype == 0x0001:
                dims = self._dataspace(body)
            elif mtype == 0x0003:
                dtype, _ = self._datatype(body)
            elif mtype == 0x0005:
                if body[0] not in (1, 2, 3):
                    raise H5Error("fill value message version")
            elif mtype == 0x0008:
                if body[0] != 3 or body[1] != 1:
                    raise H5Error("only version-3 contiguous layouts")
                layout = struct.unpack_from("<QQ", body, 2)
            elif mtype == 0x000C:
                k, v = self._attribute(body)
                attrs[k] = v
            elif mtype != 0:
                raise H5Error(f"unexpected message type {mtype:#x}")
        if dims is None or dtype is None or layout is None:
            raise H5Error("dataset lacks a dataspace, datatype or layout message")
        data_addr, nbytes = layout
        count = int(np.prod(dims))
        if nbytes != count * dtype.itemsize:
            raise H5Error("layout size disagrees with the dataspace")
        if count == 0:
            data = np.zeros(dims, dtype=dtype)
        else:
            if data_addr == UNDEF or data_addr + nbytes > len(self.buf):
                raise H5Error("dataset storage beyond the end of the file")
            data = np.frombuffer(self.buf, dtype=dtype, count=count, offset=data_addr).reshape(dims)
        return data, attrs

    # -- h5py-like access
    def keys(self):
        return sorted(self.datasets)

    def __getitem__(self, name):
        return self.datasets[name][0]

    def attrs(self, name):
        return self.datasets[name][1]
