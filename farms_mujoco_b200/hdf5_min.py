"""Minimal HDF5 writer / reader for the simulation log (``Simulation.postprocess``).

The reference saves ``task.data`` through farms_core's ``AnimatData.to_file`` -> h5py
(simulation.py:198-209).  Neither h5py nor libhdf5 exists in this image, so this module writes the
subset of the HDF5 file format the log needs, by hand, following the HDF5 File Format
Specification (version 1.1 structures, the ones every libhdf5 release reads):

* superblock version 0, 8-byte offsets and lengths;
* old-style groups: object header version 1 with a symbol-table message, one version-1 B-tree
  node pointing at one symbol-table node (``SNOD``) and a local heap with the link names
  (the group leaf-node K of the superblock is raised so that one node holds every child);
* datasets: object header version 1 with dataspace (version 1), datatype (version 1: IEEE
  little-endian floats, two's-complement integers, fixed-length null-padded ASCII strings) and
  data-layout (version 3, contiguous) messages; no filters, no chunking, no attributes.

A nested ``dict`` becomes groups, NumPy arrays / scalars / lists of ``str`` become datasets.
``read_hdf5`` parses exactly this subset back (it is the structural self-check of the writer, not
an HDF5 library); ``tests/test_hdf5_min.py`` also walks the one real HDF5 file this image holds
(a MATLAB 7.4 / HDF5 1.6 file in scipy's test data) with the same reader: superblock, root symbol
table entry, B-tree node, symbol-table node, local heap and the float64 datatype message have the
layout written here.  NOT VALIDATED AGAINST libhdf5 HERE: ``tests/test_hdf5_min.py`` reads the files
back with h5py wherever h5py is importable and is skipped otherwise; ``postprocess`` therefore keeps
writing ``simulation.npz`` next to ``simulation.hdf5``.
"""

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b'\x89HDF\r\n\x1a\n'


def _pad8(data):
    return data + b'\x00'*(-len(data) % 8)


def _message(mtype, body):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), 0) + body


def _object_header(messages):
    body = b''.join(messages)
    # version, reserved, number of messages, reference count, header size, 4 bytes to the 8-boundary
    return struct.pack('<BxHII4x', 1, len(messages), 1, len(body)) + body


def _datatype(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == 'f':
        size = dtype.itemsize
        exp_bits, man_bits = {4: (8, 23), 8: (11, 52)}[size]
        bits = bytes([0x20, size*8 - 1, 0])                  # little-endian, implied mantissa msb; sign bit position
        props = struct.pack('<HHBBBBI', 0, size*8, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        return bytes([0x11]) + bits + struct.pack('<I', size) + props
    if dtype.kind in 'iu':
        bits = bytes([0x08 if dtype.kind == 'i' else 0x00, 0, 0])   # little-endian; bit 3: signed
        return bytes([0x10]) + bits + struct.pack('<I', dtype.itemsize) + struct.pack('<HH', 0, dtype.itemsize*8)
    if dtype.kind == 'S':
        return bytes([0x13]) + bytes([0x01, 0, 0]) + struct.pack('<I', dtype.itemsize)   # null-padded ASCII
    raise TypeError(f'hdf5_min: dtype {dtype} is outside the subset')


def _as_array(value):
    if isinstance(value, (list, tuple)) and all(isinstance(v, str) for v in value):
        width = max([len(v.encode('ascii')) for v in value] + [1])
        return np.array([v.encode('ascii') for v in value], dtype=f'S{width}').reshape(len(value))
    if isinstance(value, str):
        return np.array(value.encode('ascii'), dtype=f'S{max(1, len(value))}')
    arr = np.asarray(value)
    if arr.dtype.kind == 'b':
        arr = arr.astype(np.uint8)
    if arr.dtype.kind == 'U':
        return _as_array([str(v) for v in arr.ravel().tolist()]).reshape(arr.shape)
    if arr.dtype.kind in 'fiu' and arr.dtype.byteorder == '>':
        arr = arr.astype(arr.dtype.newbyteorder('<'))
    return arr if arr.flags.c_contiguous else np.ascontiguousarray(arr)


class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data):
        self.buf += b'\x00'*(-len(self.buf) % 8)
        address = len(self.buf)
        self.buf += data
        return address

    def dataset(self, value):
        arr = _as_array(value)
        raw = arr.tobytes()
        data_address = self.alloc(raw) if raw else UNDEF
        dims = arr.shape
        space = struct.pack('<BBB5x', 1, len(dims), 0) + b''.join(struct.pack('<Q', d) for d in dims)
        layout = struct.pack('<BBQQ', 3, 1, data_address, len(raw))
        # fill-value message as libhdf5 writes it for a plain dataset (version 1: late allocation,
        # write the fill value if set, defined with size 0) -- byte for byte what a MATLAB 7.4 /
        # HDF5 1.6 file in scipy's test data carries
        fill = bytes([1, 2, 2, 1]) + struct.pack('<I', 0)
        return self.alloc(_object_header([
            _message(0x0005, fill), _message(0x0003, _datatype(arr.dtype)), _message(0x0001, space),
            _message(0x0008, layout)]))


LEAF_K = 64          # up to 128 children per group in one symbol-table node
INTERNAL_K = 16


def write_hdf5(path, tree):
    """Write the nested dict ``tree`` (groups) of arrays / scalars / lists of str (datasets)."""
    writer = _Writer()
    writer.buf += b'\x00'*96                               # the superblock goes here

    def group(node):
        names = sorted(node)
        if len(names) > 2*LEAF_K:
            raise ValueError(f'hdf5_min: more than {2*LEAF_K} children in one group')
        children = {}
        for name in names:
            value = node[name]
            children[name] = group(value)[0] if isinstance(value, dict) else writer.dataset(value)
        return _finish_group(writer, names, children)

    root_header, root_btree, root_heap = group(tree)
    writer.buf += b'\x00'*(-len(writer.buf) % 8)
    superblock = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    superblock += struct.pack('<QQQQ', 0, UNDEF, len(writer.buf), UNDEF)
    superblock += struct.pack('<QQII', 0, root_header, 1, 0) + struct.pack('<QQ', root_btree, root_heap)
    assert len(superblock) == 96
    writer.buf[0:96] = superblock
    with open(path, 'wb') as f:
        f.write(bytes(writer.buf))


def _finish_group(writer, names, children):
    """Local heap (the empty string at offset 0, then the link names), one symbol-table node with
    every child sorted by name, one B-tree node pointing at it (key 0 = "", key 1 = the largest
    name), and the group's object header (symbol-table message)."""
    heap, offsets = bytearray(8), {}
    for name in names:
        offsets[name] = len(heap)
        heap += _pad8(name.encode('ascii') + b'\x00')
    heap_data = writer.alloc(bytes(heap))
    heap_address = writer.alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), 1, heap_data))
    entries = b''.join(struct.pack('<QQII16x', offsets[name], children[name], 0, 0) for name in names)
    node = b'SNOD' + struct.pack('<BxH', 1, len(names)) + entries + b'\x00'*(40*(2*LEAF_K - len(names)))
    snod = writer.alloc(node)
    tree_node = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1 if names else 0, UNDEF, UNDEF)
    keys = struct.pack('<QQQ', 0, snod, offsets[names[-1]] if names else 0)
    tree_node += keys + b'\x00'*(8*(2*INTERNAL_K + 1) + 8*2*INTERNAL_K - len(keys))
    btree = writer.alloc(tree_node)
    # (+ an empty NIL message: libhdf5's own group headers are never shorter than 32 bytes of messages)
    header = writer.alloc(_object_header([_message(0x0011, struct.pack('<QQ', btree, heap_address)), _message(0x0000, b'')]))
    return header, btree, heap_address


# ------------------------------------------------------------------------------------ reader
def _read_messages(buf, address):
    version, n_messages, _, size = struct.unpack_from('<BxHII', buf, address)
    assert version == 1, 'object header version'
    pos, end, out = address + 16, address + 16 + size, []
    for _ in range(n_messages):
        mtype, msize, _ = struct.unpack_from('<HHB', buf, pos)
        out.append((mtype, buf[pos + 8:pos + 8 + msize]))
        pos += 8 + msize
    assert pos <= end
    return out


def _read_object(buf, address):
    messages = dict(_read_messages(buf, address))
    if 0x0011 in messages:
        btree, heap = struct.unpack_from('<QQ', messages[0x0011], 0)
        assert buf[heap:heap + 4] == b'HEAP' and buf[btree:btree + 4] == b'TREE'
        heap_data = struct.unpack_from('<Q', buf, heap + 24)[0]
        node_type, level, used = struct.unpack_from('<BBH', buf, btree + 4)
        assert node_type == 0 and level == 0
        out = {}
        for child in range(used):
            snod = struct.unpack_from('<Q', buf, btree + 24 + 8 + 16*child)[0]
            assert buf[snod:snod + 4] == b'SNOD'
            n_symbols = struct.unpack_from('<H', buf, snod + 6)[0]
            for k in range(n_symbols):
                name_offset, header = struct.unpack_from('<QQ', buf, snod + 8 + 40*k)
                start = heap_data + name_offset
                name = bytes(buf[start:buf.index(b'\x00', start)]).decode('ascii')
                out[name] = _read_object(buf, header)
        return out
    space, dtype_msg, layout = messages[0x0001], messages[0x0003], messages[0x0008]
    rank = space[1]
    dims = struct.unpack_from(f'<{rank}Q', space, 8) if rank else ()
    cls, size = dtype_msg[0] & 0x0F, struct.unpack_from('<I', dtype_msg, 4)[0]
    if cls == 1:
        dtype = np.dtype(f'<f{size}')
    elif cls == 0:
        dtype = np.dtype(f'<{"i" if dtype_msg[1] & 0x08 else "u"}{size}')
    elif cls == 3:
        dtype = np.dtype(f'S{size}')
    else:
        raise TypeError(f'datatype class {cls}')
    version, layout_class, data_address, data_size = struct.unpack_from('<BBQQ', layout, 0)
    assert version == 3 and layout_class == 1
    count = int(np.prod(dims, dtype=np.int64)) if rank else 1
    assert data_size == count*dtype.itemsize
    if data_size == 0:
        return np.zeros(dims, dtype=dtype)
    return np.frombuffer(bytes(buf[data_address:data_address + data_size]), dtype=dtype).reshape(dims).copy()


def read_hdf5(path):
    """Read back a file written by ``write_hdf5`` (this subset only) as a nested dict."""
    with open(path, 'rb') as f:
        buf = f.read()
    assert buf[:8] == SIGNATURE and buf[8] == 0 and buf[13] == 8 and buf[14] == 8
    assert struct.unpack_from('<Q', buf, 40)[0] == len(buf), 'end-of-file address'
    root_header = struct.unpack_from('<Q', buf, 56 + 8)[0]
    return _read_object(buf, root_header)
