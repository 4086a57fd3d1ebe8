"""Minimal pure-NumPy HDF5 reader (and a matching test writer) for the reference's datasets.

The reference stores its datasets with h5py / libhdf5 defaults (karel_env/generator.py:59,
132-153; vizdoom_env/generator.py; loaded by karel_env/dataset_karel.py:14-36 and
vizdoom_env/dataset_vizdoom.py): superblock version 0, old-style groups (symbol table =
v1 B-tree + local heap + SNOD nodes), version-1 object headers, contiguous (or compact /
chunked+gzip) dataset layouts, fixed-point / float / fixed-string / enum(bool) / variable-length
string datatypes.  h5py is not installed in this image, so this module reads exactly that
subset with `mmap` + `numpy.frombuffer` (contiguous datasets are zero-copy views) and exposes
the small part of the h5py API the loaders use:

    f = File(path); g = f['data_info']; g['num_train'][()]; 'env_type' in g; f[id]['s_h'][()]

Format reference: "HDF5 File Format Specification Version 2.0" (public).  Pinned against the
genuine h5py-written file the reference ships (karel_env/asset/texture.hdf5, see
tests/test_hdf5.py) and against round trips through `write_hdf5` below, which emits the same
subset (a test / conversion helper, not verified against libhdf5 itself).
"""
import mmap
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b'\x89HDF\r\n\x1a\n'


class HDF5Error(IOError):
    pass


def _u(buf, off, n):
    return int.from_bytes(buf[off:off + n], 'little')


class _Datatype(object):
    """Parsed datatype message -> numpy dtype (+ bool enum / vlen string flags)."""

    def __init__(self, buf, off):
        cv = buf[off]
        self.cls, self.version = cv & 0x0F, cv >> 4
        bits = buf[off + 1:off + 4]
        self.size = _u(buf, off + 4, 4)
        self.is_bool = False
        self.vlen_str = False
        p = off + 8
        if self.cls == 0:      # fixed point
            order = '>' if bits[0] & 1 else '<'
            signed = bool(bits[0] & 8)
            self.dtype = np.dtype('%s%s%d' % (order, 'i' if signed else 'u', self.size))
            self.length = 8 + 4
        elif self.cls == 1:    # floating point
            order = '>' if bits[0] & 1 else '<'
            self.dtype = np.dtype('%sf%d' % (order, self.size))
            self.length = 8 + 12
        elif self.cls == 3:    # fixed-length string
            self.dtype = np.dtype('S%d' % self.size)
            self.length = 8
        elif self.cls == 8:    # enumeration (h5py stores numpy bool as enum{FALSE=0, TRUE=1} over int8)
            n = bits[0] | (bits[1] << 8)
            base = _Datatype(buf, p)
            q = p + base.length
            names = []
            for _ in range(n):
                e = buf.find(b'\x00', q)
                names.append(bytes(buf[q:e]))
                ln = e - q + 1
                q += ln if self.version >= 3 else (ln + 7) // 8 * 8
            q += n * base.size
            self.dtype = base.dtype
            self.is_bool = sorted(names) == [b'FALSE', b'TRUE']
            self.length = q - off
        elif self.cls == 9:    # variable length (strings only)
            if (bits[0] & 0x0F) != 1:
                raise HDF5Error('variable-length sequences are not supported')
            self.vlen_str = True
            self.dtype = np.dtype('V%d' % self.size)
            base = _Datatype(buf, p)
            self.length = 8 + base.length
        else:
            raise HDF5Error('unsupported HDF5 datatype class %d' % self.cls)


class Dataset(object):
    def __init__(self, f, name, msgs):
        self._f, self.name = f, name
        buf = f._buf
        sp = msgs[0x0001][0]
        ver, rank = buf[sp], buf[sp + 1]
        q = sp + (8 if ver == 1 else 4)
        self.shape = tuple(_u(buf, q + 8 * i, 8) for i in range(rank))
        self._dt = _Datatype(buf, msgs[0x0003][0])
        self.dtype = np.dtype(bool) if self._dt.is_bool else self._dt.dtype
        self._layout = msgs[0x0008][0]
        self._filters = msgs.get(0x000B, [None])[0]

    def __len__(self):
        return self.shape[0]

    @property
    def size(self):
        n = 1
        for d in self.shape:
            n *= int(d)
        return n

    def _raw(self):
        buf, lo, dt = self._f._buf, self._layout, self._dt.dtype
        n = self.size
        ver = buf[lo]
        if ver == 3:
            cls = buf[lo + 1]
            if cls == 0:       # compact
                sz = _u(buf, lo + 2, 2)
                return np.frombuffer(buf, dt, n, lo + 4) if sz else np.zeros(n, dt)
            if cls == 1:       # contiguous
                addr = _u(buf, lo + 2, 8)
                if addr == UNDEF:
                    return np.zeros(n, dt)
                return np.frombuffer(buf, dt, n, addr + self._f._base)
            if cls == 2:       # chunked
                nd = buf[lo + 2]
                bt = _u(buf, lo + 3, 8)
                cdims = [_u(buf, lo + 11 + 4 * i, 4) for i in range(nd)]
                return self._read_chunks(bt, cdims[:-1])
        elif ver in (1, 2):
            nd, cls = buf[lo + 1], buf[lo + 2]
            q = lo + 8
            if cls == 1:
                addr = _u(buf, q, 8)
                return np.frombuffer(buf, dt, n, addr + self._f._base) if addr != UNDEF else np.zeros(n, dt)
            if cls == 2:
                bt = _u(buf, q, 8)
                cdims = [_u(buf, q + 8 + 4 * i, 4) for i in range(nd)]
                return self._read_chunks(bt, cdims[:-1])
            if cls == 0:
                q += 4 * nd
                return np.frombuffer(buf, dt, n, q + 4)
        raise HDF5Error('unsupported data layout (version %d)' % ver)

    def _filter_ids(self):
        if self._filters is None:
            return []
        buf, q = self._f._buf, self._filters
        ver, nf = buf[q], buf[q + 1]
        q += 8 if ver == 1 else 2
        ids = []
        for _ in range(nf):
            fid = _u(buf, q, 2)
            if ver == 1 or fid >= 256:
                nlen = _u(buf, q + 2, 2)
                ncd = _u(buf, q + 6, 2)
                q += 8 + ((nlen + 7) // 8 * 8 if ver == 1 else nlen)
            else:
                ncd = _u(buf, q + 4, 2)
                q += 6
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            ids.append(fid)
        return ids

    def _read_chunks(self, btree, cdims):
        buf, dt = self._f._buf, self._dt.dtype
        out = np.zeros(self.shape, dt)
        if btree == UNDEF:
            return out.reshape(-1)
        nd = len(cdims)
        filters = self._filter_ids()
        for f in filters:
            if f not in (1, 2):
                raise HDF5Error('unsupported HDF5 filter id %d (deflate and shuffle only)' % f)

        def walk(addr):
            a = addr + self._f._base
            if buf[a:a + 4] != b'TREE' or buf[a + 4] != 1:
                raise HDF5Error('bad chunk B-tree node')
            level, used = buf[a + 5], _u(buf, a + 6, 2)
            q = a + 24
            ksz = 8 + 8 * (nd + 1)
            for i in range(used):
                csize, mask = _u(buf, q, 4), _u(buf, q + 4, 4)
                offs = [_u(buf, q + 8 + 8 * d, 8) for d in range(nd)]
                child = _u(buf, q + ksz, 8)
                if level > 0:
                    walk(child)
                else:
                    raw = bytes(buf[child + self._f._base:child + self._f._base + csize])
                    for j, f in enumerate(reversed(filters)):
                        if mask & (1 << (len(filters) - 1 - j)):
                            continue
                        if f == 1:
                            raw = zlib.decompress(raw)
                        else:   # shuffle
                            a8 = np.frombuffer(raw, np.uint8)
                            raw = a8.reshape(dt.itemsize, -1).T.tobytes()
                    chunk = np.frombuffer(raw, dt, int(np.prod(cdims))).reshape(cdims)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, self.shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
                q += ksz + 8
        walk(btree)
        return out.reshape(-1)

    def _read(self):
        raw = self._raw()
        if self._dt.vlen_str:
            vals = [self._f._global_heap_object(bytes(r.tobytes())) for r in raw]
            arr = np.array(vals, dtype=object).reshape(self.shape)
            return arr[()] if self.shape == () else arr
        arr = raw.reshape(self.shape)
        if self._dt.is_bool:
            arr = arr.astype(bool)
        elif not arr.dtype.isnative:
            arr = arr.astype(arr.dtype.newbyteorder('='))
        return arr

    def __getitem__(self, key):
        arr = self._read()
        if key == () or key is Ellipsis:
            return arr[()] if self.shape == () else arr
        return arr[key]

    def stored(self):
        """The dataset in its STORED element type, without the bool / byte-order conversion copies of
        [()] (an enum(bool) dataset comes back as its 0/1 bytes): a zero-copy view of the memory-mapped
        file for contiguous and compact layouts.  For bulk consumers that convert while they copy."""
        if self._dt.vlen_str:
            return self._read()
        return self._raw().reshape(self.shape)

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self._read())
        return a.astype(dtype) if dtype is not None else a


class Group(object):
    def __init__(self, f, name, btree, heap):
        self._f, self.name = f, name
        self._btree, self._heap = btree, heap
        self._links = None

    def _load(self):
        if self._links is not None:
            return self._links
        f, buf = self._f, self._f._buf
        h = self._heap + f._base
        if buf[h:h + 4] != b'HEAP':
            raise HDF5Error('bad local heap')
        hdata = _u(buf, h + 24, 8) + f._base
        links = {}

        def walk(addr):
            a = addr + f._base
            if buf[a:a + 4] == b'TREE':
                if buf[a + 4] != 0:
                    raise HDF5Error('bad group B-tree node')
                used = _u(buf, a + 6, 2)
                q = a + 24 + 8
                for _ in range(used):
                    walk(_u(buf, q, 8))
                    q += 16
            elif buf[a:a + 4] == b'SNOD':
                n = _u(buf, a + 6, 2)
                q = a + 8
                for _ in range(n):
                    noff, ohdr = _u(buf, q, 8), _u(buf, q + 8, 8)
                    s = hdata + noff
                    e = buf.find(b'\x00', s)
                    links[bytes(buf[s:e]).decode('utf-8')] = ohdr
                    q += 40
            else:
                raise HDF5Error('bad symbol table node')
        if self._btree != UNDEF:
            walk(self._btree)
        self._links = links
        return links

    def keys(self):
        return list(self._load().keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, name):
        if isinstance(name, bytes):
            name = name.decode('utf-8')
        node = self
        parts = [p for p in name.split('/') if p]
        if name.startswith('/'):
            node = self._f
        for i, p in enumerate(parts):
            if not isinstance(node, Group):
                raise KeyError(name)
            links = node._load()
            if p not in links:
                raise KeyError("Unable to open object (object '%s' doesn't exist)" % p)
            node = node._f._open(links[p], (node.name.rstrip('/') + '/' + p))
        return node


class File(Group):
    """Read-only HDF5 file (the h5py.File subset the reference's loaders use)."""

    def __init__(self, path, mode='r'):
        if mode != 'r':
            raise HDF5Error('hdf5_lite.File is read-only (use write_hdf5 to create test files)')
        self._fh = open(path, 'rb')
        try:
            self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise HDF5Error('empty file: %s' % path)
        buf = self._buf
        off = 0
        while buf[off:off + 8] != SIG:      # superblock may sit at 0, 512, 1024, ...
            off = 512 if off == 0 else off * 2
            if off + 8 > len(buf):
                raise HDF5Error('%s is not an HDF5 file' % path)
        ver = buf[off + 8]
        if ver not in (0, 1):
            raise HDF5Error('HDF5 superblock version %d is not supported (h5py/libhdf5 defaults '
                            'write version 0)' % ver)
        if buf[off + 13] != 8 or buf[off + 14] != 8:
            raise HDF5Error('only 8-byte offsets/lengths are supported')
        q = off + 24 + (4 if ver == 1 else 0)
        self._base = _u(buf, q, 8)
        entry = q + 32
        ohdr, cache = _u(buf, entry + 8, 8), _u(buf, entry + 16, 4)
        self._cache = {}
        if cache == 1:
            Group.__init__(self, self, '/', _u(buf, entry + 24, 8), _u(buf, entry + 32, 8))
        else:
            root = self._open(ohdr, '/')
            Group.__init__(self, self, '/', root._btree, root._heap)
        self.filename = path

    def _messages(self, addr):
        """{type: [offset of message data, ...]} of a version-1 object header."""
        buf = self._buf
        a = addr + self._base
        if buf[a] != 1:
            raise HDF5Error('object header version %d is not supported (old-style files only)' % buf[a])
        nmsg = _u(buf, a + 2, 2)
        size = _u(buf, a + 8, 4)
        blocks = [(a + 16, size)]
        msgs = {}
        seen = 0
        while blocks and seen < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and seen < nmsg:
                mtype, msize = _u(buf, q, 2), _u(buf, q + 2, 2)
                data = q + 8
                if mtype == 0x0010:      # continuation
                    blocks.append((_u(buf, data, 8) + self._base, _u(buf, data + 8, 8)))
                else:
                    msgs.setdefault(mtype, []).append(data)
                q = data + msize
                seen += 1
        return msgs

    def _open(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        if 0x0011 in msgs:
            st = msgs[0x0011][0]
            obj = Group(self, name, _u(self._buf, st, 8), _u(self._buf, st + 8, 8))
        elif 0x0001 in msgs and 0x0003 in msgs and 0x0008 in msgs:
            obj = Dataset(self, name, msgs)
        elif 0x0002 in msgs or 0x0006 in msgs:
            raise HDF5Error('new-style (link message) groups are not supported: %s' % name)
        else:
            raise HDF5Error('unsupported HDF5 object at %d (%s)' % (addr, name))
        self._cache[addr] = obj
        return obj

    def _global_heap_object(self, ref):
        n, addr, idx = struct.unpack('<IQI', ref[:16])
        if addr == 0 and n == 0:
            return b''
        buf, a = self._buf, addr + self._base
        if buf[a:a + 4] != b'GCOL':
            raise HDF5Error('bad global heap collection')
        end = a + _u(buf, a + 8, 8)
        q = a + 16
        while q + 16 <= end:
            oi, sz = _u(buf, q, 2), _u(buf, q + 8, 8)
            if oi == idx:
                return bytes(buf[q + 16:q + 16 + n])
            if oi == 0:
                break
            q += 16 + (sz + 7) // 8 * 8
        raise HDF5Error('global heap object %d not found' % idx)

    def close(self):
        self._cache = {}
        try:
            self._buf.close()
        except (BufferError, ValueError):
            pass      # zero-copy views still alive: the map goes away with them
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ------------------------------------------------------------------------------------------------
# Writer for the same subset: nested dict of arrays / scalars / bytes -> superblock v0 file with
# old-style groups and contiguous datasets (test fixtures, synthetic -> data.hdf5 conversion).
# ------------------------------------------------------------------------------------------------
_LEAF_K, _INT_K = 4, 16               # libhdf5 defaults: <= 8 symbols per SNOD, <= 32 children per node


def _dtype_msg(dt, is_bool):
    def fixed(d):
        bits = (0 if d.byteorder in '<=|' else 1) | (8 if d.kind == 'i' else 0)
        return struct.pack('<BBBBI', 0x10, bits, 0, 0, d.itemsize) + struct.pack('<HH', 0, d.itemsize * 8)
    if is_bool:
        base = fixed(np.dtype('i1'))
        names = b'FALSE\x00\x00\x00' + b'TRUE\x00\x00\x00\x00'
        return struct.pack('<BBBBI', 0x18, 2, 0, 0, 1) + base + names + b'\x00\x01'
    if dt.kind in 'iu':
        return fixed(dt)
    if dt.kind == 'f':
        if dt.itemsize == 4:
            prop = struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31)
        elif dt.itemsize == 8:
            prop = struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63)
        else:
            raise HDF5Error('float%d not supported by the writer' % (8 * dt.itemsize))
        return struct.pack('<BBBBI', 0x11, bits[0], bits[1], 0, dt.itemsize) + prop
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0, 0, 0, dt.itemsize)
    raise HDF5Error('dtype %s not supported by the writer' % dt)


def write_hdf5(path, tree):
    """tree: {name: ndarray | scalar | bytes | str | dict (sub-group)}."""
    out = bytearray(b'\x00' * 96)      # superblock v0 (56 B) + root symbol table entry (40 B)

    def align():
        while len(out) % 8:
            out.append(0)

    def put(b):
        align()
        a = len(out)
        out.extend(b)
        return a

    def header(messages):
        body = bytearray()
        for mtype, data in messages:
            data = bytes(data) + b'\x00' * (-len(data) % 8)
            body += struct.pack('<HHBBBB', mtype, len(data), 0, 0, 0, 0) + data
        return put(struct.pack('<BBHII', 1, 0, len(messages), 1, len(body)) + b'\x00' * 4 + bytes(body))

    def dataset(value):
        if isinstance(value, str):
            value = value.encode('utf-8')
        arr = np.asarray(value)
        if arr.dtype.kind == 'U':
            arr = arr.astype('S')
        is_bool = arr.dtype == np.bool_
        raw = np.ascontiguousarray(arr.astype('i1') if is_bool else arr).tobytes()
        addr = put(raw) if raw else UNDEF
        space = struct.pack('<BBBBI', 1, arr.ndim, 0, 0, 0) + b''.join(struct.pack('<Q', d) for d in arr.shape)
        layout = struct.pack('<BBQQ', 3, 1, addr, len(raw))
        return header([(0x0001, space), (0x0003, _dtype_msg(arr.dtype, is_bool)), (0x0008, layout)])

    def group(d):
        names = sorted(d.keys(), key=lambda s: s.encode('utf-8'))
        addrs = {n: (group(d[n])[0] if isinstance(d[n], dict) else dataset(d[n])) for n in names}
        heap = bytearray(b'\x00' * 8)
        noff = {}
        for n in names:
            noff[n] = len(heap)
            heap += n.encode('utf-8') + b'\x00'
            heap += b'\x00' * (-len(heap) % 8)
        hdata = put(bytes(heap) + b'\x00' * 16)
        haddr = put(b'HEAP' + struct.pack('<BBBBQQQ', 0, 0, 0, 0, len(heap) + 16, len(heap), hdata))
        # free-list block at the end of the data segment: (next = 1 "none", size = 16)
        out[hdata + len(heap):hdata + len(heap) + 16] = struct.pack('<QQ', 1, 16)
        per = 2 * _LEAF_K
        # level-0 children: symbol nodes; (address, heap offset of the last name it holds)
        level = []
        for s in range(0, len(names), per):
            chunk = names[s:s + per]
            body = b''.join(struct.pack('<QQII', noff[n], addrs[n], 0, 0) + b'\x00' * 16 for n in chunk)
            body += b'\x00' * (40 * (per - len(chunk)))
            level.append((put(b'SNOD' + struct.pack('<BBH', 1, 0, len(chunk)) + body), noff[chunk[-1]]))
        fan, nsize = 2 * _INT_K, 24 + (2 * _INT_K) * 16 + 8
        depth = 0
        while True:     # B-tree levels, bottom-up; every node has the fixed on-disk size
            groups = [level[i:i + fan] for i in range(0, len(level), fan)] or [[]]
            align()
            base = len(out)
            nxt, left_key = [], 0
            for gi, grp in enumerate(groups):
                lsib = base + (gi - 1) * nsize if gi > 0 else UNDEF
                rsib = base + (gi + 1) * nsize if gi + 1 < len(groups) else UNDEF
                node = b'TREE' + struct.pack('<BBHQQ', 0, depth, len(grp), lsib, rsib)
                node += struct.pack('<Q', left_key)
                for child, last in grp:
                    node += struct.pack('<QQ', child, last)
                    left_key = last
                node += b'\x00' * (nsize - len(node))
                nxt.append((put(node), left_key))
            level = nxt
            depth += 1
            if len(level) == 1:
                break
        baddr = level[0][0]
        ohdr = header([(0x0011, struct.pack('<QQ', baddr, haddr))])
        return ohdr, baddr, haddr

    ohdr, baddr, haddr = group(tree)
    align()
    sb = SIG + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INT_K, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, len(out), UNDEF)
    sb += struct.pack('<QQII', 0, ohdr, 1, 0) + struct.pack('<QQ', baddr, haddr)
    out[:len(sb)] = sb
    with open(path, 'wb') as fh:
        fh.write(bytes(out))
