"""TensorFlow checkpoint (tensor bundle) reader/writer, SURVEY §8(f)-3: format constants pinned
against the published LevelDB / TensorFlow values, round trips, corruption detection.  CPU only
(d2p_crc32c is host code inside libd2p.so)."""
import os
import struct

import numpy as np
import pytest

from demo2program_b200 import tf_checkpoint as tfc


def test_crc32c_known_answers():
    # the vectors of LevelDB's / TensorFlow's own crc32c_test (from RFC 3720 B.4) + the classic check value
    assert tfc.crc32c(b'123456789') == 0xE3069283
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA
    assert tfc.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfc.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    iscsi_read = bytes([0x01, 0xc0, 0x00, 0x00] + [0] * 12 + [0x14, 0, 0, 0, 0, 0, 0x04, 0, 0, 0, 0, 0x14,
                       0, 0, 0, 0x18, 0x28, 0, 0, 0, 0, 0, 0, 0, 0x02, 0, 0, 0, 0, 0, 0, 0])
    assert tfc.crc32c(iscsi_read) == 0xD9963A56


def test_crc32c_is_incremental_and_alignment_independent():
    rs = np.random.RandomState(0)
    buf = rs.randint(0, 256, 100003).astype(np.uint8)
    whole = tfc.crc32c(buf)
    for cut in (0, 1, 7, 8, 9, 4099, 100003):
        assert tfc.crc32c(buf[cut:], tfc.crc32c(buf[:cut])) == whole
    assert tfc.crc32c(buf[3:].copy()) == tfc.crc32c(buf[3:])     # unaligned view vs aligned copy


def test_crc_mask_round_trip_and_definition():
    for c in (0, 1, 0xE3069283, 0xffffffff, 0x80000000):
        m = tfc.mask_crc(c)
        assert m == ((((c >> 15) | (c << 17)) & 0xffffffff) + 0xa282ead8) & 0xffffffff
        assert tfc.unmask_crc(m) == c
    c = tfc.crc32c(b'foo')
    assert tfc.mask_crc(c) != c and tfc.mask_crc(tfc.mask_crc(c)) != c       # as in crc32c_test "Mask"


def test_table_round_trip_many_blocks(tmp_path):
    items = [(b'', b'header')] + [(('scope%03d/var_%d' % (i // 7, i)).encode(), os.urandom(i % 97)) for i in range(900)]
    items = sorted(set(items))
    p = str(tmp_path / 't.index')
    tfc.write_table(p, items, block_size=1024)
    assert tfc.read_table(p) == items
    raw = open(p, 'rb').read()
    assert struct.unpack('<Q', raw[-8:])[0] == 0xdb4775248b80fb57 and len(raw) > 48
    with pytest.raises(ValueError):
        tfc.write_table(p, [(b'b', b''), (b'a', b'')])


def test_snappy_blocks_are_readable():
    # literal "abcd", copy(offset 4, len 8) with overlap, 2-byte-offset copy, long literal
    body = bytes([3 << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4]) + bytes([((3 - 1) << 2) | 2, 12, 0]) + \
        bytes([60 << 2, 69]) + bytes(range(70))
    exp = b'abcd' + b'abcdabcd' + b'abc' + bytes(range(70))
    assert tfc._snappy_uncompress(bytes([len(exp)]) + body) == exp
    with pytest.raises(tfc.CheckpointError):
        tfc._snappy_uncompress(bytes([5]) + bytes([3 << 2]) + b'abcd')


def _variables():
    rs = np.random.RandomState(1)
    return {
        'Demo_Encoder/State_Encoder/conv1/Conv/weights': rs.randn(3, 3, 16, 16).astype(np.float32),
        'Demo_Encoder/State_Encoder/conv1/Conv/biases': np.zeros(16, np.float32),
        'Program_Decoder/Token_Embedding/embedding_map': rs.randn(51, 512).astype(np.float32),
        'global_step': np.int64(1234),
        'optimizer_pixel_loss/beta1_power': np.float32(0.9 ** 5),
        'some/int32': np.arange(-3, 9, dtype=np.int32).reshape(3, 4),
        'some/bool': np.array([True, False, True]),
        'some/f64': rs.randn(2, 2),
        'some/empty': np.zeros((0, 4), np.float32),
    }


def test_bundle_round_trip_and_layout(tmp_path):
    v = _variables()
    prefix = str(tmp_path / 'train_dir' / 'model-1234')
    tfc.save_checkpoint(prefix, v)
    assert sorted(os.listdir(tmp_path / 'train_dir')) == ['checkpoint', 'model-1234.data-00000-of-00001',
                                                          'model-1234.index']
    got = tfc.load_checkpoint(prefix)
    assert sorted(got) == sorted(v)
    for k in v:
        a = np.asarray(v[k])
        assert got[k].dtype == a.dtype and got[k].shape == a.shape and np.array_equal(got[k], a), k
    # tensors lie in the data file in byte-wise key order, back to back
    entries = dict(tfc.read_table(prefix + '.index'))
    assert tfc._decode_header(entries[b'']) == {'num_shards': 1, 'endianness': 0}
    off = 0
    for name in sorted(v, key=str.encode):
        e = tfc._decode_entry(entries[name.encode()])
        assert e['offset'] == off and e['size'] == np.asarray(v[name]).nbytes and e['shard_id'] == 0
        off += e['size']
    assert os.path.getsize(prefix + '.data-00000-of-00001') == off
    assert tfc.list_variables(prefix)[0][0] == 'Demo_Encoder/State_Encoder/conv1/Conv/biases'
    sub = tfc.load_checkpoint(prefix, names=['global_step'])
    assert list(sub) == ['global_step'] and int(sub['global_step']) == 1234
    with pytest.raises(KeyError):
        tfc.load_checkpoint(prefix, names=['nope'])


def test_entry_proto_bytes_are_the_documented_wire_format():
    # BundleEntryProto{dtype: DT_FLOAT(1), shape{dim{size:3} dim{size:4}}, offset: 300, size: 48, crc32c: fixed32}
    b = tfc._encode_entry(np.float32, (3, 4), 300, 48, 0x01020304)
    assert b == bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x03, 0x12, 0x02, 0x08, 0x04,
                       0x20, 0xac, 0x02, 0x28, 0x30, 0x35, 0x04, 0x03, 0x02, 0x01])
    e = tfc._decode_entry(b)
    assert (e['dtype'], e['shape'], e['offset'], e['size'], e['crc32c']) == (1, [3, 4], 300, 48, 0x01020304)
    assert tfc._encode_header(1) == bytes([0x08, 0x01, 0x1a, 0x02, 0x08, 0x01])


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / 'model-1')
    tfc.save_checkpoint(prefix, _variables(), update_state=False)
    data = prefix + '.data-00000-of-00001'
    raw = bytearray(open(data, 'rb').read())
    raw[100] ^= 0x40
    open(data, 'wb').write(raw)
    with pytest.raises(tfc.CheckpointError, match='tensor checksum'):
        tfc.load_checkpoint(prefix)
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[10] ^= 1
    open(prefix + '.index', 'wb').write(idx)
    with pytest.raises(tfc.CheckpointError, match='block checksum'):
        tfc.load_checkpoint(prefix)
    open(prefix + '.index', 'wb').write(b'x' * 100)
    with pytest.raises(tfc.CheckpointError, match='magic'):
        tfc.load_checkpoint(prefix)


def test_checkpoint_state_file(tmp_path):
    d = str(tmp_path)
    assert tfc.latest_checkpoint(d) is None
    for step in (0, 1000, 2000):
        tfc.save_checkpoint(os.path.join(d, 'model-%d' % step), {'global_step': np.int64(step)})
    txt = open(os.path.join(d, 'checkpoint')).read().splitlines()
    assert txt[0] == 'model_checkpoint_path: "model-2000"'
    assert txt[1:] == ['all_model_checkpoint_paths: "model-%d"' % s for s in (0, 1000, 2000)]
    assert tfc.latest_checkpoint(d) == os.path.join(d, 'model-2000')
    assert tfc.is_tf_checkpoint(os.path.join(d, 'model-1000')) and not tfc.is_tf_checkpoint(os.path.join(d, 'model-5'))


def test_optimizer_slot_names_follow_tf_slot_naming():
    state = {'a/kernel': np.ones((2, 2), np.float32), 'a/BatchNorm/moving_mean': np.zeros(2, np.float32)}
    m = {'a/kernel': np.full((2, 2), 0.5, np.float32)}
    v = {'a/kernel': np.full((2, 2), 0.25, np.float32)}
    out = tfc.with_optimizer_slots(state, m, v, step=3)
    assert set(out) == {'a/kernel', 'a/BatchNorm/moving_mean', 'optimizer_pixel_loss/a/kernel/Adam',
                        'optimizer_pixel_loss/a/kernel/Adam_1', 'optimizer_pixel_loss/beta1_power',
                        'optimizer_pixel_loss/beta2_power', 'optimizer_pixel_loss/learning_rate',
                        'global_step'}
    assert out['global_step'].dtype == np.int64
    assert np.isclose(out['optimizer_pixel_loss/beta1_power'], 0.9 ** 4)
    st, m2, v2, step = tfc.split_optimizer_slots(out)
    assert step == 3 and set(m2) == {'a/kernel'} and np.array_equal(v2['a/kernel'], v['a/kernel'])
    assert set(st) == {'a/kernel', 'a/BatchNorm/moving_mean', 'global_step'}


def test_full_model_variable_set_round_trips_by_name(tmp_path):
    """Every variable of the Karel `full` manifest (SURVEY Appendix B names) + slots, 135 MB."""
    from demo2program_b200.config import karel_config
    from demo2program_b200.manifest import build_manifests
    pm, sm = build_manifests(karel_config('full', batch_size=32, k=10))
    flat, sflat = pm.init_flat(0), sm.init_flat(0)
    state = {e.name: pm.view(flat, e.name) for e in pm}
    state.update({e.name: sm.view(sflat, e.name) for e in sm})
    m = {e.name: pm.view(flat, e.name) * 0.1 for e in pm}
    v = {e.name: pm.view(flat, e.name) ** 2 for e in pm}
    prefix = str(tmp_path / 'model-7')
    tfc.save_checkpoint(prefix, tfc.with_optimizer_slots(state, m, v, 7))
    names = [n for n, _, _ in tfc.list_variables(prefix)]
    assert len(names) == 3 * len(pm.entries) + len(sm.entries) + 4      # + beta powers, learning_rate, global_step
    st, m2, v2, step = tfc.split_optimizer_slots(tfc.load_checkpoint(prefix))
    assert step == 7
    for e in pm:
        assert np.array_equal(st[e.name], state[e.name]) and np.array_equal(v2[e.name], v[e.name])
