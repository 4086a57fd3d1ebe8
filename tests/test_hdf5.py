"""HDF5 reader (demo2program_b200/hdf5_lite.py): the genuine h5py-written asset of the reference,
round trips through the test writer, and the dataset surface on top of it (reference
karel_env/dataset_karel.py:14-160, input_ops_karel.py:52-116)."""
import hashlib
import json
import os

import numpy as np
import pytest

from demo2program_b200 import hdf5_lite
from demo2program_b200.config import karel_config
from demo2program_b200.synthetic import make_batch

HERE = os.path.dirname(os.path.abspath(__file__))
ASSET = '/root/reference/karel_env/asset/texture.hdf5'
GOLDEN = os.path.join(HERE, 'golden', 'texture_hdf5.json')


def _digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def test_golden_fixture_is_committed():
    g = json.load(open(GOLDEN))
    assert set(g) >= {'wall', 'marker'}
    for v in g.values():
        assert set(v) == {'shape', 'dtype', 'sha256', 'sum', 'min', 'max'}


@pytest.mark.skipif(not os.path.exists(ASSET), reason='reference asset not mounted (GPU box)')
def test_reads_the_reference_asset_written_by_h5py():
    """texture.hdf5 is the only HDF5 file the reference ships (superblock v0, symbol-table groups,
    contiguous datasets): names, shapes, dtypes and content hashes are pinned in tests/golden/."""
    g = json.load(open(GOLDEN))
    with hdf5_lite.File(ASSET) as f:
        assert sorted(f.keys()) == sorted(g.keys())
        for name, want in g.items():
            d = f[name]
            assert list(d.shape) == want['shape'] and str(d.dtype) == want['dtype']
            a = d[()]
            assert _digest(a) == want['sha256']
            assert abs(float(a.sum()) - want['sum']) < 1e-6 * max(1.0, abs(want['sum']))
            assert float(a.min()) == want['min'] and float(a.max()) == want['max']


def test_round_trip_of_every_supported_type(tmp_path):
    rs = np.random.RandomState(0)
    tree = {
        'b': rs.rand(3, 4, 5) > 0.5,
        'i8': rs.randint(-100, 100, (7,)).astype(np.int8), 'u8': rs.randint(0, 255, (2, 9)).astype(np.uint8),
        'i16': rs.randint(-3000, 3000, (4, 3)).astype(np.int16), 'i32': rs.randint(-10 ** 6, 10 ** 6, (5,)).astype(np.int32),
        'i64': np.arange(6, dtype=np.int64).reshape(2, 3) * 10 ** 12, 'f32': rs.randn(8, 2).astype(np.float32),
        'f64': rs.randn(3, 3), 'scalar_int': 42, 'scalar_float': 2.5, 'text': 'prob', 'empty': np.zeros((0, 4), np.float32),
        'sub': {'x': np.arange(10), 'deeper': {'y': np.float32(1.5)}},
    }
    path = str(tmp_path / 'rt.hdf5')
    hdf5_lite.write_hdf5(path, tree)
    with hdf5_lite.File(path) as f:
        assert sorted(f.keys()) == sorted(tree.keys())
        for k in ('b', 'i8', 'u8', 'i16', 'i32', 'i64', 'f32', 'f64', 'empty'):
            a = f[k][()]
            assert a.dtype == tree[k].dtype and a.shape == tree[k].shape and np.array_equal(a, tree[k]), k
        assert int(f['scalar_int'][()]) == 42 and float(f['scalar_float'][()]) == 2.5
        assert f['text'][()] == b'prob'
        assert np.array_equal(f['sub']['x'][()], np.arange(10)) and np.array_equal(f['sub/x'][2:5], [2, 3, 4])
        assert float(f['sub/deeper/y'][()]) == 1.5
        assert 'sub' in f and 'nope' not in f and 'deeper' in f['sub']
        with pytest.raises(KeyError):
            f['nope']


def test_group_btree_with_many_links(tmp_path):
    """A dataset file has one group per example in the root group: several B-tree levels."""
    n = 3000
    tree = {'ex_%05d' % i: {'v': np.array([i, i * i], np.int64)} for i in range(n)}
    path = str(tmp_path / 'many.hdf5')
    hdf5_lite.write_hdf5(path, tree)
    with hdf5_lite.File(path) as f:
        assert len(f) == n
        for i in (0, 1, 7, 8, 255, 256, 257, 1999, n - 1):
            assert f['ex_%05d' % i]['v'][()].tolist() == [i, i * i]


def test_rejects_non_hdf5_and_new_style_files(tmp_path):
    p = tmp_path / 'x.hdf5'
    p.write_bytes(b'not an hdf5 file' * 100)
    with pytest.raises(hdf5_lite.HDF5Error):
        hdf5_lite.File(str(p))
    p.write_bytes(hdf5_lite.SIG + bytes([2]) + b'\x00' * 200)     # superblock version 2
    with pytest.raises(hdf5_lite.HDF5Error):
        hdf5_lite.File(str(p))


def test_dataset_directory_round_trip(tmp_path):
    """synthetic examples -> data.hdf5 + id.txt in the generator's schema -> the reference's loader
    logic (get_data: padding to max_demo_len, one-hot actions with quirk F10, --num_k prefix) ->
    the same feed-dict arrays the synthetic generator produces directly."""
    from demo2program_b200 import dataset as ds
    k, test_k, num_k = 5, 3, 4
    d = str(tmp_path / 'karel_ds')
    ids = ds.write_karel_dataset(d, 6, 3, 2, k, test_k=test_k, seed=40)
    tr, te, va = ds.create_default_splits(d, num_k=num_k)
    assert (len(tr), len(te), len(va)) == (6, 3, 2)
    assert sorted(tr.ids + te.ids + va.ids) == sorted(ids)
    assert tr.dsl_type == 'prob' and tr.max_demo_len == 20 and tr.num_action_tokens == 5
    cfg = karel_config('full', batch_size=1, k=k)
    cfg.test_k = test_k
    for split in (tr, te, va):
        for ex_id in split.ids:
            i = int(ex_id.split('_')[1])
            src = make_batch(cfg, seed=40 + i, batch_size=1)
            got = ds.collate(split, [ex_id])
            for key in ds.KEYS:
                want = src[key]
                if key in ('s_h', 'a_h', 'a_h_tokens', 'demo_len', 'per'):
                    want = want[:, :num_k]          # --num_k selects a prefix of the seen demos
                assert got[key].shape == want.shape, (key, got[key].shape, want.shape)
                assert np.array_equal(got[key], want.astype(got[key].dtype)), key
            assert got['s_h'].dtype == np.uint8 and got['program_tokens'].dtype == np.int32


def test_batches_iterator_over_hdf5_dataset(tmp_path):
    from demo2program_b200 import dataset as ds
    d = str(tmp_path / 'karel_ds2')
    ds.write_karel_dataset(d, 8, 2, 2, 3, test_k=2, seed=7)
    tr, _, _ = ds.create_default_splits(d, num_k=3)
    it = ds.batches(tr, 4, shuffle=True, seed=1, epochs=1)
    seen = []
    for b in it:
        assert b['s_h'].shape == (4, 3, 20, 8, 8, 16) and b['program'].shape == (4, 50, 50)
        seen += [x.decode() for x in b['id']]
    assert sorted(seen) == sorted(tr.ids)


def test_loader_processes_deliver_the_same_batches_in_the_same_order(tmp_path):
    from demo2program_b200 import dataset as ds
    d = str(tmp_path / 'karel_ds3')
    ds.write_karel_dataset(d, 12, 2, 2, 3, test_k=2, seed=3)
    tr, _, _ = ds.create_default_splits(d, num_k=2)
    a = list(ds.batches(tr, 4, shuffle=True, seed=5, epochs=2))
    b = list(ds.batches(tr, 4, shuffle=True, seed=5, epochs=2, workers=2))
    assert len(a) == len(b) == 6
    for x, y in zip(a, b):
        assert set(x) == set(y)
        for key in x:
            assert x[key].dtype == y[key].dtype and np.array_equal(x[key], y[key]), key


def test_vizdoom_dataset_directory(tmp_path):
    """vizdoom_env/dataset_vizdoom.py:14-140: --num_k slices the stored seen demos before padding,
    actions carry quirk F10, init positions are padded to vizdoom_max_init_pos_len."""
    from demo2program_b200 import dataset as ds
    rs = np.random.RandomState(4)
    K, TK, T, L, V, A, P, PL = 5, 3, 7, 12, 20, 11, 6, 4
    tree, ids, src = {}, [], {}
    for i in range(6):
        n = rs.randint(4, L + 1)
        lens, tlens = rs.randint(2, T + 1, K), rs.randint(2, T + 1, TK)
        m, tm = int(lens.max()), int(tlens.max())
        ex = {'program': rs.randint(0, V, n).astype(np.int8),
              's_h': rs.randint(0, 256, (K, m, 6, 5, 3)).astype(np.int16),
              'test_s_h': rs.randint(0, 256, (TK, tm, 6, 5, 3)).astype(np.int16),
              'a_h': rs.randint(0, A, (K, m - 1)).astype(np.int8), 'test_a_h': rs.randint(0, A, (TK, tm - 1)).astype(np.int8),
              's_h_len': lens.astype(np.int16), 'test_s_h_len': tlens.astype(np.int16),
              'p_v_h': rs.rand(K, m, P) > 0.5, 'test_p_v_h': rs.rand(TK, tm, P) > 0.5,
              'vizdoom_init_pos': rs.randn(K, 2, 3, 2), 'vizdoom_init_pos_len': rs.randint(0, 4, (K, 2)).astype(np.int16),
              'test_vizdoom_init_pos': rs.randn(TK, 2, 2, 2), 'test_vizdoom_init_pos_len': rs.randint(0, 3, (TK, 2)).astype(np.int16)}
        name = 'viz_%d' % i
        tree[name], src[name] = ex, ex
        ids.append(name)
    tree['data_info'] = {'num_demo_per_program': np.int64(K), 'num_test_demo_per_program': np.int64(TK),
                         'max_demo_length': np.int64(T), 'max_program_length': np.int64(L),
                         'num_program_tokens': np.int64(V), 'num_action_tokens': np.int64(A),
                         'vizdoom_pos_keys': np.array([b'player_pos', b'demon_pos']), 'vizdoom_max_init_pos_len': np.int64(PL),
                         'perception_type': 'simple', 'level': 'simple', 's_h_h': np.int64(6), 's_h_w': np.int64(5),
                         's_h_c': np.int64(3), 'num_train': np.int64(4), 'num_test': np.int64(1), 'num_val': np.int64(1)}
    d = tmp_path / 'viz'
    d.mkdir()
    hdf5_lite.write_hdf5(str(d / 'data.hdf5'), tree)
    (d / 'id.txt').write_text('\n'.join(ids) + '\n')
    num_k = 3
    tr, te, va = ds.create_default_splits(str(d), num_k=num_k, dataset_type='vizdoom')
    assert isinstance(tr, ds.H5DatasetVizdoom) and (len(tr), len(te), len(va)) == (4, 1, 1)
    assert tr.vizdoom_pos_keys == ['player_pos', 'demon_pos'] and tr.level == 'simple' and tr.test_k == TK
    for ex_id in tr.ids:
        e = src[ex_id]
        t = tr.get_data(ex_id)
        assert len(t) == 17
        program, ptok, s_h, ts_h, a_h, a_tok, ta_h, ta_tok, plen, dlen, tdlen, per, tper, ip, ipl, tip, tipl = t
        n = len(e['program'])
        assert program.shape == (V, L) and program.sum() == n and ptok[:n].tolist() == e['program'].tolist()
        assert s_h.shape == (num_k, T, 6, 5, 3) and ts_h.shape == (TK, T, 6, 5, 3)
        m = e['s_h'].shape[1]
        assert np.array_equal(s_h[:, :m], e['s_h'][:num_k]) and not s_h[:, m:].any()
        assert a_h.shape == (num_k, T, A + 1) and a_h[:, m - 1, A].all()        # <e> at the program-max position (F10)
        assert np.array_equal(a_tok[:, :m - 1], e['a_h'][:num_k])
        assert dlen.tolist() == e['s_h_len'][:num_k].tolist() and tdlen.tolist() == e['test_s_h_len'].tolist()
        assert per.shape == (num_k, T, P) and np.array_equal(per[:, :m], e['p_v_h'][:num_k])
        assert ip.shape == (num_k, 2, PL, 2) and np.array_equal(ip[:, :, :3], e['vizdoom_init_pos'][:num_k])
        assert not ip[:, :, 3:].any() and tip.shape == (TK, 2, PL, 2)
        assert ipl.tolist() == e['vizdoom_init_pos_len'][:num_k].tolist()
    b = ds.collate(tr, tr.ids[:2])
    assert b['s_h'].dtype == np.uint8 and b['init_pos'].shape == (2, num_k, 2, PL, 2)
    assert set(ds.VIZDOOM_EXTRA_KEYS) <= set(b)
    _same_batches(b, ds.collate(tr, tr.ids[:2], fast=False))      # direct fill == get_data + stack + astype


def _same_batches(x, y):
    assert set(x) == set(y)
    for key in x:
        assert x[key].dtype == y[key].dtype and x[key].shape == y[key].shape, key
        assert np.array_equal(x[key], y[key]), key


def test_direct_fill_collate_equals_get_data_stack(tmp_path):
    """collate()'s default path for the HDF5 datasets (_fast_collate: batch arrays filled straight from the
    memory-mapped file, stored-array views cached per example) against the general path over the
    golden-pinned get_data: same keys, shapes, dtypes and bytes - first touch and cached, every --num_k."""
    from demo2program_b200 import dataset as ds
    d = str(tmp_path / 'karel_fast')
    ds.write_karel_dataset(d, 10, 3, 2, 5, test_k=3, seed=11)
    for num_k in (5, 2, 7):            # 7 > stored demos: the prefix is everything stored
        tr, te, _ = ds.create_default_splits(d, num_k=num_k)
        for split in (tr, te):
            for rep in range(2):       # second pass: cached views
                for lst in (split.ids[:3], split.ids[1:2], list(reversed(split.ids))):
                    _same_batches(ds.collate(split, lst), ds.collate(split, lst, fast=False))
        assert tr._stored_views and len(tr._stored_views) <= len(tr.ids)


def test_loader_processes_without_the_parent_copy(tmp_path):
    """batches(..., workers=2, copy=False): the same batches in the same order when every batch is consumed
    on receipt (what the trainer does), and a yielded batch stays intact while ONE more batch is requested
    (its shared-memory slot is recycled only with the second further request)."""
    from demo2program_b200 import dataset as ds
    d = str(tmp_path / 'karel_nocopy')
    ds.write_karel_dataset(d, 24, 2, 2, 3, test_k=2, seed=9)
    tr, _, _ = ds.create_default_splits(d, num_k=3)
    want = list(ds.batches(tr, 4, shuffle=True, seed=2, epochs=2))
    got, prev = [], None
    for b in ds.batches(tr, 4, shuffle=True, seed=2, epochs=2, workers=2, copy=False, lookahead=2):
        snap = {k: np.array(v, copy=True) for k, v in b.items()}
        if prev is not None:           # the previous batch's windows are still what they were
            for k in prev[0]:
                assert np.array_equal(prev[0][k], prev[1][k]), k
        prev = (b, snap)
        got.append(snap)
    assert len(got) == len(want) == 12
    for x, y in zip(got, want):
        _same_batches(x, y)


def test_karel_loader_matches_reference_loader_golden(tmp_path):
    """tests/golden/dataset_karel_golden.json: digests of the 13-tuples the REFERENCE's
    karel_env/dataset_karel.py Dataset.get_data returns (run in the build container by
    tests/golden/make_dataset_golden.py over hdf5_lite-as-h5py) on a seeded dataset directory, and
    the reference's shuffled split ids.  Our loader must return the same arrays (shape, dtype,
    bytes) in the same split order."""
    from demo2program_b200 import dataset as ds
    g = json.load(open(os.path.join(HERE, 'golden', 'dataset_karel_golden.json')))
    sp = g['spec']
    d = str(tmp_path / 'golden_ds')
    ds.write_karel_dataset(d, sp['n_train'], sp['n_test'], sp['n_val'], sp['k'], test_k=sp['test_k'], seed=sp['seed'])
    ds.rs = np.random.RandomState(123)          # the module-level shuffler of a fresh process
    tr, te, va = ds.create_default_splits(d, num_k=g['num_k'])
    assert list(tr.ids) == g['splits']['train'] and list(te.ids) == g['splits']['test'] and list(va.ids) == g['splits']['val']
    for k_, v in g['attrs'].items():
        assert getattr(tr, k_) == v
    n = 0
    for split in (tr, te, va):
        for ex_id in split.ids:
            got = split.get_data(ex_id)
            want = g['examples'][ex_id]
            assert len(got) == len(want) == 13
            for j, (a, w) in enumerate(zip(got, want)):
                a = np.ascontiguousarray(np.asarray(a))
                assert list(a.shape) == w['shape'], (ex_id, j, a.shape, w['shape'])
                assert str(a.dtype) == w['dtype'], (ex_id, j, a.dtype, w['dtype'])
                assert _digest(a) == w['sha256'], (ex_id, j)
            n += 1
    assert n == 17


def test_action_one_hots_equal_the_reference_loop():
    """dataset._action_one_hots (vectorised) against the per-demonstration loop of
    karel_env/dataset_karel.py:66-77, incl. quirk F10 (the end token sits behind the zero PADDED
    row, padding zeros count as action 0) and the IndexError when the stored rows are T long."""
    import pytest
    from demo2program_b200.dataset import _action_one_hots
    rs = np.random.RandomState(3)
    T, A = 20, 5
    for n, m in [(10, 7), (1, 1), (5, 19), (3, 0)]:
        a = rs.randint(0, A, size=(n, m)).astype(np.int64)
        a[:, m // 2:] *= rs.randint(0, 2, size=(n, m - m // 2))       # zero padding inside the rows
        hist = []
        for t in a:
            h = np.zeros([T, A + 1], dtype=bool)
            h[np.arange(len(t)), t] = 1
            h[len(t), A] = 1
            hist.append(h)
        ref = np.stack(hist, 0)
        got, tok = _action_one_hots(a, T, A)
        assert got.dtype == np.bool_ and np.array_equal(got, ref)
        assert np.array_equal(tok, np.argmax(ref, axis=2))
    with pytest.raises(IndexError):
        _action_one_hots(np.zeros((2, T), np.int64), T, A)


def test_rank_shards_are_disjoint_and_cover_the_split(tmp_path):
    """batches(rank=r, world=W): rank r draws from ids[r::W] - no example twice per epoch across the ranks,
    every example in exactly one shard; a shard smaller than the batch size fails at the first next()."""
    import pytest
    from demo2program_b200 import dataset as ds
    d = str(tmp_path / 'karel_shards')
    ds.write_karel_dataset(d, 16, 2, 2, 3, test_k=2, seed=21)
    tr, _, _ = ds.create_default_splits(d, num_k=3)
    seen = []
    for r in range(2):
        mine = []
        for b in ds.batches(tr, 4, shuffle=True, seed=r, epochs=1, rank=r, world=2):
            mine += [x.decode() for x in b['id']]
        assert len(mine) == 8 and len(set(mine)) == 8 and set(mine) == set(tr.ids[r::2])
        seen += mine
    assert sorted(seen) == sorted(tr.ids)
    with pytest.raises(ValueError, match='fewer than batch_size'):
        next(ds.batches(tr, 4, rank=0, world=8))        # 2 examples per shard


def test_stored_views_are_zero_copy_windows_of_the_file(tmp_path):
    """hdf5_lite.Dataset.stored(): enum(bool) datasets come back as their 0/1 bytes (int8, h5py's enum base type), integers in their stored
    width, both as windows of the memory map (no copy) for contiguous layouts; [()] converts as before."""
    d = tmp_path / 'v.h5'
    frames = np.random.RandomState(0).rand(3, 4, 5) > 0.5
    hdf5_lite.write_hdf5(str(d), {'g': {'b': frames, 'i': np.arange(6, dtype=np.int16).reshape(2, 3)}})
    f = hdf5_lite.File(str(d))
    sb, si = f['g']['b'].stored(), f['g']['i'].stored()
    assert sb.dtype == np.int8 and sb.shape == frames.shape and np.array_equal(sb, frames.astype(np.int8))   # h5py's enum base
    assert si.dtype == np.int16 and np.array_equal(si, np.arange(6).reshape(2, 3))
    assert not sb.flags.owndata and not sb.flags.writeable and not si.flags.owndata
    assert f['g']['b'][()].dtype == np.bool_ and np.array_equal(f['g']['b'][()], frames)
