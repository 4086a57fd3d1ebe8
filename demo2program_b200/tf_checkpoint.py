"""TensorFlow-1.x checkpoint files (the "tensor bundle", Saver write_version V2) by variable name.

SURVEY §8(f) row 3.  The reference saves with `tf.train.Saver(max_to_keep=100)`
(trainer.py:114,182-186) and restores with `pretrain_saver.restore` (trainer.py:145: trainable
variables only) / `saver.restore` (evaler.py:82-99, `tf.train.latest_checkpoint` at evaler.py:86).
TF 1.3 cannot be installed here, so this module reads and writes the on-disk format itself:

  <prefix>.index                an SSTable (the LevelDB table format TF vendors under
                                tensorflow/core/lib/io): sorted string keys -> protobuf values;
                                key "" -> BundleHeaderProto, every other key = variable name ->
                                BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
  <prefix>.data-00000-of-00001  the raw little-endian tensor bytes, concatenated in key order
  checkpoint                    CheckpointState text proto (`model_checkpoint_path: "..."`)

Every SSTable block carries a 5-byte trailer (compression type + masked CRC-32C), every tensor a
masked CRC-32C in its entry; both are verified on read and produced on write (the CRC runs in
libd2p.so, `d2p_crc32c`).  Blocks may be stored raw or snappy-compressed (both are read; written
raw, which is what TF's BundleWriter does).  Partitioned ("sliced") variables are not used by the
reference and are rejected.

The variable names are exactly those of `manifest.py`; optimizer slots follow
TF's slot naming under `optimize_loss(name='optimizer_pixel_loss')` (trainer.py:102-109):
`optimizer_pixel_loss/<variable>/Adam`, `.../Adam_1`, `optimizer_pixel_loss/beta{1,2}_power`,
plus `global_step` (int64).

Parity status: the container format is restated from the published LevelDB table / tensor-bundle
layout and pinned against their published constants (table magic, CRC-32C known answers from
LevelDB's own crc32c test, masking rule); no TensorFlow-written file exists offline to read back,
so cross-reading with real TF is **unpinned**.
"""
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
BLOCK_TRAILER = 5
RESTART_INTERVAL = 16
BLOCK_SIZE = 256 * 1024
MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DT = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'),
       5: np.dtype('<i2'), 6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?'),
       17: np.dtype('<u2'), 19: np.dtype('<f2'), 22: np.dtype('<u4'), 23: np.dtype('<u8')}
_DT_OF = {v: k for k, v in _DT.items()}


class CheckpointError(IOError):
    pass


# ----------------------------------------------------------------------------- checksums
_CRC_TABLE = None


def _crc32c_numpy(buf, crc=0):
    """CRC-32C (Castagnoli, reflected 0x82F63B78) without libd2p.so: table-driven, one byte at a time
    (slow - checkpoints can still be read on a host where the CUDA library cannot be loaded)."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.arange(256, dtype=np.uint64)
        for _ in range(8):
            t = np.where(t & 1, (t >> np.uint64(1)) ^ np.uint64(0x82F63B78), t >> np.uint64(1))
        _CRC_TABLE = [int(x) for x in t]
    c = (crc ^ 0xffffffff) & 0xffffffff
    tab = _CRC_TABLE
    for b in buf.tobytes():
        c = tab[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def crc32c(data, crc=0):
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else \
        np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    if buf.size == 0:
        return crc
    try:
        from . import _lib
        lib = _lib.load()
    except Exception:       # no libd2p.so / no CUDA driver on this host: pure-Python fallback
        return _crc32c_numpy(buf, crc)
    return int(lib.d2p_crc32c(buf.ctypes.data, buf.size, crc))


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + MASK_DELTA) & 0xffffffff


def unmask_crc(m):
    rot = (m - MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ----------------------------------------------------------------------------- varints / protobuf
def _put_varint(out, v):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift = v = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError('truncated varint')
        b = buf[pos]
        pos += 1
        v |= (b & 0x7f) << shift
        if b < 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise CheckpointError('varint too long')


def _pb_fields(buf):
    """Yields (field number, wire type, value) of one serialized message."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        fno, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            if len(v) != n:
                raise CheckpointError('truncated protobuf field')
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError('unsupported protobuf wire type %d' % wt)
        yield fno, wt, v


def _pb_varint_field(out, fno, v):
    _put_varint(out, fno << 3)
    _put_varint(out, v)


def _pb_bytes_field(out, fno, payload):
    _put_varint(out, (fno << 3) | 2)
    _put_varint(out, len(payload))
    out.extend(payload)


def _encode_entry(dtype, shape, offset, size, crc_masked):
    """BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto)."""
    out = bytearray()
    _pb_varint_field(out, 1, _DT_OF[np.dtype(dtype)])
    sh = bytearray()
    for d in shape:
        dim = bytearray()
        _pb_varint_field(dim, 1, int(d))
        _pb_bytes_field(sh, 2, dim)
    _pb_bytes_field(out, 2, sh)
    if offset:
        _pb_varint_field(out, 4, offset)
    _pb_varint_field(out, 5, size)
    _put_varint(out, (6 << 3) | 5)
    out.extend(struct.pack('<I', crc_masked))
    return bytes(out)


def _decode_entry(buf):
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': 0, 'slices': 0}
    for fno, wt, v in _pb_fields(buf):
        if fno == 1:
            e['dtype'] = v
        elif fno == 2:
            for f2, _, dimbuf in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, dv in _pb_fields(dimbuf):
                        if f3 == 1:
                            size = dv
                    e['shape'].append(size)
                elif f2 == 3 and dimbuf:
                    raise CheckpointError('tensor of unknown rank in checkpoint')
        elif fno == 3:
            e['shard_id'] = v
        elif fno == 4:
            e['offset'] = v
        elif fno == 5:
            e['size'] = v
        elif fno == 6:
            e['crc32c'] = v
        elif fno == 7:
            e['slices'] += 1
    return e


def _encode_header(num_shards=1):
    """BundleHeaderProto: num_shards, endianness LITTLE (default 0), version {producer: 1}."""
    out = bytearray()
    _pb_varint_field(out, 1, num_shards)
    ver = bytearray()
    _pb_varint_field(ver, 1, 1)
    _pb_bytes_field(out, 3, ver)
    return bytes(out)


def _decode_header(buf):
    h = {'num_shards': 1, 'endianness': 0}
    for fno, wt, v in _pb_fields(buf):
        if fno == 1:
            h['num_shards'] = v
        elif fno == 2:
            h['endianness'] = v
    return h


# ----------------------------------------------------------------------------- snappy (read only)
def _snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError('corrupt snappy block')
        start = len(out) - off
        for i in range(ln):            # copies may overlap their own output
            out.append(out[start + i])
    if len(out) != n:
        raise CheckpointError('snappy block has %d bytes, header says %d' % (len(out), n))
    return bytes(out)


# ----------------------------------------------------------------------------- SSTable
def _read_block(data, offset, size):
    end = offset + size + BLOCK_TRAILER
    if end > len(data):
        raise CheckpointError('block handle beyond end of file')
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    stored = struct.unpack_from('<I', data, offset + size + 1)[0]
    if unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
        raise CheckpointError('block checksum mismatch at offset %d' % offset)
    if ctype == 0:
        return bytes(contents)
    if ctype == 1:
        return _snappy_uncompress(bytes(contents))
    raise CheckpointError('unknown block compression type %d' % ctype)


def _block_entries(block):
    """(key, value) pairs of one table block (prefix-compressed keys + restart array)."""
    if len(block) < 4:
        raise CheckpointError('block too small')
    nrestart = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * nrestart
    if limit < 0:
        raise CheckpointError('bad restart array')
    pos, key = 0, b''
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > limit:
            raise CheckpointError('corrupt block entry')
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, 'rb') as f:
        data = f.read()
    if len(data) < FOOTER_LEN:
        raise CheckpointError('%s: too short for an SSTable' % path)
    footer = data[-FOOTER_LEN:]
    if struct.unpack_from('<Q', footer, 40)[0] != TABLE_MAGIC:
        raise CheckpointError('%s: not an SSTable (bad magic number)' % path)
    _, p = _get_varint(footer, 0)          # metaindex handle (unused)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        out.extend(_block_entries(_read_block(data, boff, bsize)))
    return out


class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b''

    def add(self, key, value):
        shared = 0
        if self.count % RESTART_INTERVAL == 0:
            if self.count:
                self.restarts.append(len(self.buf))
        else:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        _put_varint(self.buf, shared)
        _put_varint(self.buf, len(key) - shared)
        _put_varint(self.buf, len(value))
        self.buf += key[shared:]
        self.buf += value
        self.last = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        out = bytes(self.buf)
        out += b''.join(struct.pack('<I', r) for r in self.restarts)
        return out + struct.pack('<I', len(self.restarts))


def write_table(path, items, block_size=BLOCK_SIZE):
    """Writes sorted (key, value) byte pairs as an uncompressed SSTable."""
    items = list(items)
    for (a, _), (b, _) in zip(items, items[1:]):
        if not a < b:
            raise ValueError('table keys must be strictly increasing')
    f = bytearray()

    def emit(contents):
        off = len(f)
        f.extend(contents)
        f.append(0)                                            # kNoCompression
        f.extend(struct.pack('<I', mask_crc(crc32c(bytes(contents) + b'\x00'))))
        h = bytearray()
        _put_varint(h, off)
        _put_varint(h, len(contents))
        return bytes(h)

    index = _BlockBuilder()
    blk = _BlockBuilder()
    for key, value in items:
        blk.add(key, value)
        if blk.size() >= block_size:
            index.add(blk.last, emit(blk.finish()))            # separator = last key of the block
            blk = _BlockBuilder()
    if blk.count:
        index.add(blk.last, emit(blk.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    f.extend(footer)
    with open(path, 'wb') as fh:
        fh.write(f)


# ----------------------------------------------------------------------------- bundles
def _shard_name(prefix, shard, num):
    return '%s.data-%05d-of-%05d' % (prefix, shard, num)


def list_variables(prefix):
    """[(name, shape, numpy dtype)] of a checkpoint, without reading tensor data."""
    out = []
    for key, value in read_table(prefix + '.index'):
        if key == b'':
            continue
        e = _decode_entry(value)
        out.append((key.decode(), tuple(e['shape']), _DT.get(e['dtype'])))
    return out


def load_checkpoint(prefix, names=None, verify=True):
    """name -> ndarray for every (or the named) variable of `<prefix>.index/.data-*`."""
    items = read_table(prefix + '.index')
    if not items or items[0][0] != b'':
        raise CheckpointError('%s.index: missing bundle header' % prefix)
    hdr = _decode_header(items[0][1])
    if hdr['endianness'] != 0:
        raise CheckpointError('big-endian checkpoint')
    shards = {}
    out = {}
    want = None if names is None else set(names)
    for key, value in items[1:]:
        name = key.decode()
        if want is not None and name not in want:
            continue
        e = _decode_entry(value)
        if e['slices']:
            raise CheckpointError('%s: partitioned variables are not supported' % name)
        if e['dtype'] not in _DT:
            raise CheckpointError('%s: unsupported dtype enum %d' % (name, e['dtype']))
        dt = _DT[e['dtype']]
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap(_shard_name(prefix, sid, hdr['num_shards']), dtype=np.uint8, mode='r')
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        n = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if raw.size != e['size'] or n * dt.itemsize != e['size']:
            raise CheckpointError('%s: size mismatch (entry %d bytes, shape %s)' % (name, e['size'], e['shape']))
        raw = np.array(raw)
        if verify and unmask_crc(e['crc32c']) != crc32c(raw):
            raise CheckpointError('%s: tensor checksum mismatch' % name)
        out[name] = raw.view(dt).reshape(e['shape'])
    if want is not None and want - set(out):
        raise KeyError('not in checkpoint: %s' % sorted(want - set(out)))
    return out


def save_checkpoint(prefix, tensors, update_state=True):
    """Writes `<prefix>.index` + `<prefix>.data-00000-of-00001` (one shard), names sorted as
    the Saver does, and (update_state) the directory's `checkpoint` state file."""
    d = os.path.dirname(os.path.abspath(prefix))
    os.makedirs(d, exist_ok=True)
    items = [(b'', _encode_header(1))]
    offset = 0
    with open(_shard_name(prefix, 0, 1) + '.tmp', 'wb') as f:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name])
            dt = a.dtype.newbyteorder('<') if a.dtype.itemsize > 1 else a.dtype
            if np.dtype(dt) not in _DT_OF:
                raise TypeError('%s: dtype %s has no TF checkpoint encoding here' % (name, a.dtype))
            shape = a.shape                                  # () for scalars such as global_step
            raw = np.ascontiguousarray(a.astype(dt, copy=False)).reshape(-1).view(np.uint8)
            f.write(raw.tobytes())
            items.append((name.encode(), _encode_entry(dt, shape, offset, raw.size, mask_crc(crc32c(raw)))))
            offset += raw.size
    os.replace(_shard_name(prefix, 0, 1) + '.tmp', _shard_name(prefix, 0, 1))
    write_table(prefix + '.index.tmp', items)
    os.replace(prefix + '.index.tmp', prefix + '.index')
    if update_state:
        update_checkpoint_state(d, os.path.basename(prefix))


def update_checkpoint_state(train_dir, name):
    """The `checkpoint` file tf.train.Saver maintains (CheckpointState text proto)."""
    path = os.path.join(train_dir, 'checkpoint')
    allp = []
    if os.path.exists(path):
        allp = re.findall(r'^all_model_checkpoint_paths:\s*"(.*)"\s*$', open(path).read(), re.M)
    if name in allp:
        allp.remove(name)
    allp.append(name)
    with open(path, 'w') as f:
        f.write('model_checkpoint_path: "%s"\n' % name)
        for p in allp:
            f.write('all_model_checkpoint_paths: "%s"\n' % p)


def latest_checkpoint(train_dir):
    """tf.train.latest_checkpoint (evaler.py:86): prefix named by the state file, or None."""
    path = os.path.join(train_dir, 'checkpoint')
    if not os.path.exists(path):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', open(path).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    p = p if os.path.isabs(p) else os.path.join(train_dir, p)
    return p if os.path.exists(p + '.index') else None


def is_tf_checkpoint(path):
    return os.path.exists(path + '.index')


# ----------------------------------------------------------------------------- model <-> variables
OPT_SCOPE = 'optimizer_pixel_loss'      # optimize_loss(name=...) at reference trainer.py:108


def with_optimizer_slots(state, adam_m, adam_v, step, beta1=0.9, beta2=0.999, trainable=None, learning_rate=1e-3):
    """Adds what the reference's full Saver also stores: Adam slots, beta powers, the
    `optimizer_pixel_loss/learning_rate` variable that `optimize_loss` creates for a float learning rate
    (trainer.py:102-109; a full-graph `saver.restore` would raise NotFound without it), global_step.

    state: name -> array (model.state_dict()); adam_m / adam_v: name -> array for the trainable
    variables.  TF keeps beta^t as variables that start at beta and are multiplied after every
    step, so after `step` updates they hold beta^(step+1)."""
    out = dict(state)
    for name in (trainable if trainable is not None else adam_m):
        out['%s/%s/Adam' % (OPT_SCOPE, name)] = adam_m[name]
        out['%s/%s/Adam_1' % (OPT_SCOPE, name)] = adam_v[name]
    out[OPT_SCOPE + '/beta1_power'] = np.float32(beta1 ** (step + 1))
    out[OPT_SCOPE + '/beta2_power'] = np.float32(beta2 ** (step + 1))
    out[OPT_SCOPE + '/learning_rate'] = np.float32(learning_rate)
    out['global_step'] = np.int64(step)
    return out


def split_optimizer_slots(variables):
    """Inverse of with_optimizer_slots: (model variables, adam_m, adam_v, global_step or None)."""
    state, m, v = {}, {}, {}
    pre = OPT_SCOPE + '/'
    for name, a in variables.items():
        if name.startswith(pre):
            inner = name[len(pre):]
            if inner == 'learning_rate':
                continue
            if inner.endswith('/Adam_1'):
                v[inner[:-7]] = a
            elif inner.endswith('/Adam'):
                m[inner[:-5]] = a
        else:
            state[name] = a
    step = int(state['global_step']) if 'global_step' in state else None
    return state, m, v, step


def _by_name(manifest, flat):
    return {e.name: flat[e.offset:e.offset + e.size].reshape(e.shape) for e in manifest}


def save_model(prefix, model, include_optimizer=True):
    """What `self.saver.save(session, train_dir/model, global_step)` stores (trainer.py:182-186):
    every model variable and BatchNorm moving statistic, `global_step`, and (include_optimizer)
    the Adam slots - from the facade's flat buffers."""
    eng = model.engine
    state = dict(model.state_dict())
    state.pop('global_step', None)
    step = eng.step_count() if hasattr(eng, 'step_count') else 0
    if include_optimizer and hasattr(eng, 'adam_m'):
        m = _by_name(eng.pm, eng.adam_m.cpu().numpy())
        v = _by_name(eng.pm, eng.adam_v.cpu().numpy())
        variables = with_optimizer_slots(state, m, v, step, learning_rate=float(getattr(eng, 'lr', 1e-3)))
    else:
        variables = dict(state)
        variables['global_step'] = np.int64(step)
    save_checkpoint(prefix, variables)
    return sorted(variables)


def load_model(prefix, model, trainable_only=False, restore_optimizer=True):
    """`saver.restore` (evaler.py:99) / `pretrain_saver.restore` (trainer.py:145, trainable_only).
    Raises KeyError if a model variable is missing from the checkpoint, as Saver.restore does."""
    import torch
    eng = model.engine
    variables = load_checkpoint(prefix)
    state, m, v, step = split_optimizer_slots(variables)
    needed = [e.name for e in eng.pm] + ([] if trainable_only else [e.name for e in eng.sm])
    missing = [n for n in needed if n not in state]
    if missing:
        raise KeyError('%s: variables not found in checkpoint: %s' % (prefix, missing[:5]))
    for e in list(eng.pm) + ([] if trainable_only else list(eng.sm)):
        if tuple(state[e.name].shape) != e.shape:
            raise ValueError('%s: checkpoint shape %s, model shape %s' % (e.name, state[e.name].shape, e.shape))
    model.load_state_dict(state, trainable_only=trainable_only)
    if not trainable_only and restore_optimizer and hasattr(eng, 'adam_m') and m and v:
        fm, fv = eng.adam_m.cpu().numpy(), eng.adam_v.cpu().numpy()
        for e in eng.pm:
            if e.name in m and e.name in v:
                fm[e.offset:e.offset + e.size] = np.asarray(m[e.name], np.float32).reshape(-1)
                fv[e.offset:e.offset + e.size] = np.asarray(v[e.name], np.float32).reshape(-1)
        eng.adam_m.copy_(torch.from_numpy(fm))
        eng.adam_v.copy_(torch.from_numpy(fv))
        if step is not None:
            eng.adam_state[0] = float(step)
    return step
