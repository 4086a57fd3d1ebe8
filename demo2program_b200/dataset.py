"""Dataset surface of the reference (karel_env/dataset_karel.py,
vizdoom_env/dataset_vizdoom.py, */input_ops_*.py) for the CLI clones.

`create_default_splits(path, num_k)` returns (train, test, val) objects with
`.ids`, `.get_data(id)` (the 13-tuple of dataset_karel.py:38-115) and the
`dsl_type/env_type/...` attributes the drivers read.  Two backends:
  * `synthetic[:N]` as dataset_path -> seeded synthetic examples (no dataset is
    available offline);
  * a directory with data.hdf5 + id.txt (the generator's output, karel_env/generator.py:
    59-153) -> h5py when it is installed, else the package's own pure-NumPy reader
    (hdf5_lite.py: superblock v0 / symbol-table groups / contiguous or chunked datasets,
    zero-copy views of the memory-mapped file).
`batches(dataset, batch_size, shuffle)` replaces the TF queue pipeline
(input_ops_karel.py:24-125) with a deterministic host iterator yielding the
feed-dict dicts of models/model_full.py:185-206.
"""
import os
import os.path as osp

import numpy as np

from .config import karel_config
from .synthetic import make_batch


def open_hdf5(path):
    """h5py.File(path, 'r') when h5py is importable, else hdf5_lite.File(path)."""
    try:
        import h5py
        return h5py.File(path, 'r')
    except ImportError:
        from . import hdf5_lite
        return hdf5_lite.File(path)


def _text(v):
    v = v.tolist() if isinstance(v, np.ndarray) else v
    return v.decode('utf-8') if isinstance(v, bytes) else str(v)

rs = np.random.RandomState(123)   # reference dataset_karel.py:11


class SyntheticDataset(object):
    def __init__(self, name, n, num_k, seed, dataset_type='karel'):
        self.name, self.num_k = name, num_k
        self._ids = ['%s_%06d' % (name, i) for i in range(n)]
        self._seed = seed
        self.dsl_type, self.env_type = 'prob', None
        self._cfg = karel_config('full', batch_size=1, k=num_k)

    @property
    def ids(self):
        return self._ids

    def __len__(self):
        return len(self._ids)

    def _example(self, id):
        idx = self._ids.index(id) if isinstance(id, str) else int(id)
        return make_batch(self._cfg, seed=self._seed + idx, batch_size=1)

    def get_data(self, id):
        b = self._example(id)
        sq = lambda k: b[k][0]
        return (sq('program').astype(bool), sq('program_tokens'), sq('s_h'), sq('test_s_h'),
                sq('a_h').astype(bool), sq('a_h_tokens'), sq('test_a_h').astype(bool),
                sq('test_a_h_tokens'), sq('program_len'), sq('demo_len'), sq('test_demo_len'),
                sq('per'), sq('test_per'))


def _action_one_hots(a, T, A):
    """The reference's per-demonstration loop (karel_env/dataset_karel.py:66-77,
    vizdoom_env/dataset_vizdoom.py:86-99) for all demonstrations at once: row i of the stored
    (per-program zero-padded) action matrix `a` gives h[i, t, a[i, t]] = 1 for every stored column t,
    then the end token h[i, len(row), A] = 1; tokens = argmax.  Like the loop, it raises IndexError
    when the stored rows are already T long."""
    a = np.asarray(a)
    if a.ndim != 2:                      # ragged / empty input: keep the row-by-row form
        hist = []
        for t in a:
            h = np.zeros([T, A + 1], dtype=bool)
            h[np.arange(len(t)), t] = 1
            h[len(t), A] = 1
            hist.append(h)
        hist = np.stack(hist, 0)
        return hist, np.argmax(hist, axis=2)
    n, m = a.shape
    hist = np.zeros([n, T, A + 1], dtype=bool)
    hist[np.arange(n)[:, None], np.arange(m)[None, :], a] = 1
    hist[:, m, A] = 1
    return hist, np.argmax(hist, axis=2)


class H5Dataset(object):
    """reference karel_env/dataset_karel.py:14-115 over h5py (lazy import)."""

    def __init__(self, ids, dataset_path, name='default', num_k=10, is_train=True):
        self._ids, self.name, self.num_k = list(ids), name, num_k
        self.data = open_hdf5(osp.join(dataset_path, 'data.hdf5'))
        info = self.data['data_info']
        g = lambda k: info[k][()]
        self.dsl_type = _text(g('dsl_type'))
        self.max_demo_len = int(g('max_demo_length'))
        self.max_program_len = int(g('max_program_length'))
        self.num_program_tokens = int(g('num_program_tokens'))
        self.num_action_tokens = int(g('num_action_tokens'))
        self.env_type = _text(g('env_type')) if 'env_type' in info else None

    @property
    def ids(self):
        return self._ids

    def __len__(self):
        return len(self._ids)

    def get_data(self, id):
        d = self.data[id]
        toks = d['program'][()]
        L, T, A = self.max_program_len, self.max_demo_len, self.num_action_tokens
        program = np.zeros([self.num_program_tokens, L], dtype=bool)
        program[toks, np.arange(len(toks))] = 1
        ptoks = np.zeros([L], dtype=toks.dtype)
        ptoks[:len(toks)] = toks

        def pad_demo(x):
            out = np.zeros((x.shape[0], T) + x.shape[2:], dtype=x.dtype)
            out[:, :x.shape[1]] = x
            return out

        def actions(a):   # quirk F10: one-hots from the per-program zero-padded matrix
            return _action_one_hots(a, T, A)

        demo, tdemo = pad_demo(d['s_h'][()]), pad_demo(d['test_s_h'][()])
        ah, aht = actions(d['a_h'][()])
        tah, taht = actions(d['test_a_h'][()])
        pk, tpk = ('p_v_h', 'test_p_v_h') if 'p_v_h' in d else ('per', 'test_per')
        per, tper = pad_demo(d[pk][()]), pad_demo(d[tpk][()])
        k = self.num_k
        return (program, ptoks, demo[:k], tdemo, ah[:k], aht[:k], tah, taht,
                np.array([len(toks)], dtype=np.float32), d['s_h_len'][()][:k], d['test_s_h_len'][()],
                per[:k], tper)


class H5DatasetVizdoom(object):
    """reference vizdoom_env/dataset_vizdoom.py:14-140 over h5py or hdf5_lite: the Karel tuple plus
    (init_pos, init_pos_len, test_init_pos, test_init_pos_len); `--num_k` slices the stored seen
    demos BEFORE padding (:61, :74, :105-106, :120-125)."""

    def __init__(self, ids, dataset_path, name='default', num_k=10, is_train=True):
        self._ids, self.name, self.num_k = list(ids), name, num_k
        self.data = open_hdf5(osp.join(dataset_path, 'data.hdf5'))
        info = self.data['data_info']
        g = lambda k: info[k][()]
        self.num_demo = int(g('num_demo_per_program'))
        self.max_demo_len = int(g('max_demo_length'))
        self.max_program_len = int(g('max_program_length'))
        self.num_program_tokens = int(g('num_program_tokens'))
        self.num_action_tokens = int(g('num_action_tokens'))
        self.vizdoom_pos_keys = [_text(v) for v in np.asarray(g('vizdoom_pos_keys')).reshape(-1)]
        self.vizdoom_max_init_pos_len = int(g('vizdoom_max_init_pos_len'))
        self.perception_type = _text(g('perception_type'))
        self.level = _text(g('level')) if 'level' in info else 'not_simple'
        self.k = int(g('num_demo_per_program'))
        self.test_k = int(g('num_test_demo_per_program'))
        self.s_h_h, self.s_h_w, self.s_h_c = int(g('s_h_h')), int(g('s_h_w')), int(g('s_h_c'))
        self.dsl_type, self.env_type = 'vizdoom', None

    @property
    def ids(self):
        return self._ids

    def __len__(self):
        return len(self._ids)

    def get_data(self, id):
        d, k = self.data[id], self.num_k
        T, A, L = self.max_demo_len, self.num_action_tokens, self.max_program_len
        toks = d['program'][()]
        program = np.zeros([self.num_program_tokens, L], dtype=bool)
        program[toks, np.arange(len(toks))] = 1
        ptoks = np.zeros([L], dtype=toks.dtype)
        ptoks[:len(toks)] = toks

        def pad_time(x):
            out = np.zeros((x.shape[0], T) + x.shape[2:], dtype=x.dtype)
            out[:, :x.shape[1]] = x
            return out

        def actions(a):
            return _action_one_hots(a, T, A)

        def pad_pos(x):
            out = np.zeros([x.shape[0], x.shape[1], self.vizdoom_max_init_pos_len, 2], dtype=x.dtype)
            out[:, :, :x.shape[2], :] = x
            return out

        demo, tdemo = pad_time(d['s_h'][()][:k]), pad_time(d['test_s_h'][()])
        ah, aht = actions(d['a_h'][()][:k])
        tah, taht = actions(d['test_a_h'][()])
        per, tper = pad_time(d['p_v_h'][()][:k]), pad_time(d['test_p_v_h'][()])
        return (program, ptoks, demo, tdemo, ah, aht, tah, taht,
                np.array([len(toks)], dtype=np.float32), d['s_h_len'][()][:k], d['test_s_h_len'][()],
                per, tper, pad_pos(d['vizdoom_init_pos'][()][:k]), d['vizdoom_init_pos_len'][()][:k],
                pad_pos(d['test_vizdoom_init_pos'][()]), d['test_vizdoom_init_pos_len'][()])


def create_default_splits(dataset_path, num_k=10, is_train=True, dataset_type='karel'):
    """reference karel_env/dataset_karel.py:131-160 / vizdoom_env/dataset_vizdoom.py:157-185."""
    if str(dataset_path).startswith('synthetic'):
        n = int(dataset_path.split(':')[1]) if ':' in dataset_path else 512
        return (SyntheticDataset('train', n, num_k, 1000), SyntheticDataset('test', max(n // 8, 32), num_k, 500000),
                SyntheticDataset('val', max(n // 8, 32), num_k, 900000))
    f = open_hdf5(osp.join(dataset_path, 'data.hdf5'))
    nt, nte, nv = (int(f['data_info'][k][()]) for k in ('num_train', 'num_test', 'num_val'))
    vizdoom = dataset_type == 'vizdoom' or 'vizdoom_pos_keys' in f['data_info']
    f.close()
    with open(osp.join(dataset_path, 'id.txt')) as fp:
        ids = [s.strip() for s in fp.readlines() if s]
    tr, te, va = ids[:nt], ids[nt:nt + nte], ids[nt + nte:nt + nte + nv]
    rs.shuffle(tr); rs.shuffle(te); rs.shuffle(va)
    cls = H5DatasetVizdoom if vizdoom else H5Dataset
    mk = lambda i, n: cls(i, dataset_path, name=n, num_k=num_k, is_train=is_train)
    return mk(tr, 'train'), mk(te, 'test'), mk(va, 'val')


def write_karel_dataset(dirname, n_train, n_test, n_val, k, test_k=None, seed=0, cfg=None):
    """Write `n_train + n_test + n_val` seeded synthetic examples as a dataset directory in the
    generator's on-disk schema (data.hdf5 + id.txt; reference karel_env/generator.py:104-153 and
    the perception / unseen-demo fields karel_env/dataset_karel.py:38-58 reads): per example a
    group `no_<i>_prog_len_<n>_max_s_h_len_<m>` with `program` (int8 tokens), `s_h` / `test_s_h`
    (bool [demos, max len in the program, h, w, 16]), `a_h` / `test_a_h` (int8, zero-padded to the
    program's longest demo), `s_h_len` / `test_s_h_len`, `p_v_h` / `test_p_v_h`; plus `data_info`.
    Uses hdf5_lite.write_hdf5 (this image has no h5py).  Returns the example ids."""
    from .hdf5_lite import write_hdf5
    cfg = cfg or karel_config('full', batch_size=1, k=k)
    if test_k is not None:
        cfg.test_k = test_k
    tree, ids = {}, []
    for i in range(n_train + n_test + n_val):
        b = make_batch(cfg, seed=seed + i, batch_size=1)
        n = int(b['program_len'][0, 0])
        ex = {'program': b['program_tokens'][0, :n].astype(np.int8)}
        m_all = 0
        for pre in ('', 'test_'):
            lens = b[pre + 'demo_len'][0].astype(np.int16)
            m = int(lens.max())
            m_all = max(m_all, m)
            ex[pre + 's_h'] = b[pre + 's_h'][0, :, :m].astype(bool)
            ex[pre + 'a_h'] = b[pre + 'a_h_tokens'][0, :, :m - 1].astype(np.int8)
            ex[pre + 's_h_len'] = lens
            ex[pre + 'a_h_len'] = (lens - 1).astype(np.int16)
            ex[pre + 'p_v_h'] = b[pre + 'per'][0, :, :m].astype(bool)
        name = 'no_%d_prog_len_%d_max_s_h_len_%d' % (i, n, m_all)
        tree[name] = ex
        ids.append(name)
    tree['data_info'] = {
        'max_demo_length': np.int64(cfg.max_demo_len), 'dsl_type': 'prob',
        'max_program_length': np.int64(cfg.max_program_len),
        'num_program_tokens': np.int64(cfg.dim_program_token),
        'num_demo_per_program': np.int64(cfg.k + cfg.test_k),
        'num_action_tokens': np.int64(cfg.action_space - 1),
        'num_train': np.int64(n_train), 'num_test': np.int64(n_test), 'num_val': np.int64(n_val)}
    os.makedirs(dirname, exist_ok=True)
    write_hdf5(osp.join(dirname, 'data.hdf5'), tree)
    with open(osp.join(dirname, 'id.txt'), 'w') as fp:
        fp.write('\n'.join(ids) + '\n')
    return ids


KEYS = ('program', 'program_tokens', 's_h', 'test_s_h', 'a_h', 'a_h_tokens', 'test_a_h',
        'test_a_h_tokens', 'program_len', 'demo_len', 'test_demo_len', 'per', 'test_per')


VIZDOOM_EXTRA_KEYS = ('init_pos', 'init_pos_len', 'test_init_pos', 'test_init_pos_len')


_LOADER_TIMEOUT_S = 600         # a forked loader process that hangs fails the run instead of stalling it
_VIEW_CACHE_MAX = 1 << 18      # examples whose stored-array views are kept per dataset object (~1 KB each)


def _stored(dset):
    """A dataset's array without conversion copies where the reader offers it (hdf5_lite), else [()]."""
    f = getattr(dset, 'stored', None)
    return f() if f is not None else dset[()]


def _fast_collate(dataset, ids, alloc=None):
    """collate() for the two HDF5 dataset classes without the per-example temporaries: the batch arrays
    are allocated once in their final dtypes and every stored array is copied (and converted) straight
    from the memory-mapped file into its slot - the same values as get_data + np.stack + astype
    (tests/test_hdf5.py compares the two byte for byte), ~4x the examples per second of one loader
    thread.  Returns None when the batch needs the general path (ragged action rows).
    alloc(key, shape, dtype) -> zero-filled array: where the batch arrays live (the loader processes
    hand out windows of their shared-memory slot, so nothing is copied on the way to the parent)."""
    viz = isinstance(dataset, H5DatasetVizdoom)
    B, k = len(ids), dataset.num_k
    L, T, A, V = (dataset.max_program_len, dataset.max_demo_len, dataset.num_action_tokens,
                  dataset.num_program_tokens)
    data = dataset.data
    d0 = data[ids[0]]
    pk, tpk = ('p_v_h', 'test_p_v_h') if 'p_v_h' in d0 else ('per', 'test_per')
    s0, ts0, p0 = d0['s_h'].shape, d0['test_s_h'].shape, d0[pk].shape
    kk, tk = min(k, s0[0]), ts0[0]
    f32, i32, u8 = np.float32, np.int32, np.uint8
    spec = [('program', (B, V, L), f32), ('program_tokens', (B, L), i32),
            ('s_h', (B, kk, T) + tuple(s0[2:]), u8), ('test_s_h', (B, tk, T) + tuple(ts0[2:]), u8),
            ('a_h', (B, kk, T, A + 1), f32), ('a_h_tokens', (B, kk, T), i32),
            ('test_a_h', (B, tk, T, A + 1), f32), ('test_a_h_tokens', (B, tk, T), i32),
            ('program_len', (B, 1), f32), ('demo_len', (B, kk), f32), ('test_demo_len', (B, tk), f32),
            ('per', (B, kk, T) + tuple(p0[2:]), f32), ('test_per', (B, tk, T) + tuple(p0[2:]), f32)]
    if viz:
        PL = dataset.vizdoom_max_init_pos_len
        ip0, tip0 = d0['vizdoom_init_pos'].shape, d0['test_vizdoom_init_pos'].shape
        spec += [('init_pos', (B, kk, ip0[1], PL, 2), f32),
                 ('init_pos_len', (B, kk) + tuple(d0['vizdoom_init_pos_len'].shape[1:]), f32),
                 ('test_init_pos', (B, tk, tip0[1], PL, 2), f32),
                 ('test_init_pos_len', (B, tk) + tuple(d0['test_vizdoom_init_pos_len'].shape[1:]), f32)]
    out = {'id': np.array([str(i).encode() for i in ids])}
    for key, shape, dt in spec:
        out[key] = np.zeros(shape, dt) if alloc is None else alloc(key, shape, dt)

    def padded(dst, x, n):         # dst[:n, :t] = x[:n]; a longer stored array fails like the reference's pad
        x = x[:n]
        if x.shape[1] > dst.shape[1] or x.shape[0] != dst.shape[0]:
            raise ValueError('could not broadcast input array from shape %s into shape %s' % (x.shape, dst.shape))
        dst[:, :x.shape[1]] = x

    def one_hots(hist, tok, a, n):
        # _action_one_hots on the stored (per-program zero-padded) matrix, rows [:n] (quirk F10)
        a = np.asarray(a)
        if a.ndim != 2:
            return False
        a = a[:n]
        m = a.shape[1]
        rows = np.arange(a.shape[0])[:, None]
        hist[rows, np.arange(m)[None, :], a] = 1
        hist[:, m, A] = 1          # IndexError when the stored rows are already T long, as in the reference
        tok[:, :m] = a
        tok[:, m] = A
        return True

    names = ['program', 's_h', 'test_s_h', 'a_h', 'test_a_h', 's_h_len', 'test_s_h_len', pk, tpk]
    if viz:
        names += ['vizdoom_init_pos', 'vizdoom_init_pos_len', 'test_vizdoom_init_pos', 'test_vizdoom_init_pos_len']
    # views of an example's stored arrays (zero-copy windows of the memory-mapped file with hdf5_lite) are kept
    # from the second epoch on: what remains per example is the copies themselves
    cache = dataset.__dict__.setdefault('_stored_views', {})
    for b, ex in enumerate(ids):
        v = cache.get(ex)
        if v is None:
            d = data[ex]
            v = tuple(_stored(d[nm]) for nm in names)
            if len(cache) < _VIEW_CACHE_MAX and all(getattr(d[nm], 'stored', None) is not None for nm in names[:1]):
                cache[ex] = v
        toks = np.asarray(v[0])
        n = len(toks)
        out['program'][b, toks, np.arange(n)] = 1
        out['program_tokens'][b, :n] = toks
        out['program_len'][b, 0] = n
        padded(out['s_h'][b], v[1], kk)
        padded(out['test_s_h'][b], v[2], tk)
        if not one_hots(out['a_h'][b], out['a_h_tokens'][b], v[3], kk):
            return None
        if not one_hots(out['test_a_h'][b], out['test_a_h_tokens'][b], v[4], tk):
            return None
        out['demo_len'][b] = v[5][:kk]
        out['test_demo_len'][b] = v[6]
        padded(out['per'][b], v[7], kk)
        padded(out['test_per'][b], v[8], tk)
        if viz:
            x = v[9][:kk]
            out['init_pos'][b][:, :, :x.shape[2], :] = x
            out['init_pos_len'][b] = v[10][:kk]
            x = v[11]
            out['test_init_pos'][b][:, :, :x.shape[2], :] = x
            out['test_init_pos_len'][b] = v[12]
    return out


def collate(dataset, ids, fast=True):
    """load_fn + batch stacking (reference karel_env/input_ops_karel.py:52-116).  Frames stay
    uint8 (the on-disk bool); everything else uses the reference's dtypes.  fast: the HDF5 dataset
    classes fill the batch arrays directly (_fast_collate); False forces get_data + stack."""
    if fast and ids and type(dataset) in (H5Dataset, H5DatasetVizdoom):
        out = _fast_collate(dataset, ids)
        if out is not None:
            return out
    cols = [dataset.get_data(i) for i in ids]
    out = {'id': np.array([str(i).encode() for i in ids])}
    keys = KEYS + (VIZDOOM_EXTRA_KEYS if cols and len(cols[0]) == len(KEYS) + 4 else ())
    for j, key in enumerate(keys):
        arr = np.stack([c[j] for c in cols])
        if key in ('s_h', 'test_s_h'):
            # the stored bool frames ARE the u8 frames (0/1 bytes): reinterpret, do not copy
            arr = arr.view(np.uint8) if arr.dtype == np.bool_ else arr.astype(np.uint8)
        elif key.endswith('_tokens'):
            arr = arr.astype(np.int32)
        else:
            arr = arr.astype(np.float32)
        out[key] = arr
    return out


_WORKER_DATASET = None      # inherited by forked loader processes
_WORKER_SLOTS = None        # anonymous shared mappings the workers write the large arrays into
_BIG_KEYS = ('s_h', 'test_s_h', 'program', 'a_h', 'test_a_h', 'per', 'test_per')


def _collate_job(job):
    slot, id_list = job
    buf = _WORKER_SLOTS[slot]
    meta = {}
    state = {'off': 0}

    def alloc(key, shape, dt):     # large arrays are BUILT in the shared slot: nothing to copy afterwards
        if key not in _BIG_KEYS:
            return np.zeros(shape, dt)
        n = int(np.prod(shape)) * np.dtype(dt).itemsize
        off = state['off']
        if off + n > buf.size:
            return np.zeros(shape, dt)
        buf[off:off + n] = 0
        meta[key] = (off, tuple(shape), np.dtype(dt).str)
        state['off'] = off + (n + 63) // 64 * 64
        return buf[off:off + n].view(dt).reshape(shape)

    b = None
    if type(_WORKER_DATASET) in (H5Dataset, H5DatasetVizdoom):
        b = _fast_collate(_WORKER_DATASET, id_list, alloc)
    if b is None:
        meta.clear()
        state['off'] = 0
        b = collate(_WORKER_DATASET, id_list, fast=False)
    off = state['off']
    for key in _BIG_KEYS:          # large arrays travel through shared memory, not the result pipe
        if key in meta:
            b.pop(key)
            continue
        a = np.ascontiguousarray(b.pop(key))
        n = a.nbytes
        buf[off:off + n] = a.reshape(-1).view(np.uint8)
        meta[key] = (off, a.shape, a.dtype.str)
        off += (n + 63) // 64 * 64
    return slot, meta, b


def batches(dataset, batch_size, shuffle=True, seed=0, epochs=None, workers=0, lookahead=4, rank=0, world=1,
            copy=True):
    """Deterministic replacement of string_input_producer + shuffle_batch (reference
    karel_env/input_ops_karel.py:24-125 uses 16 loader threads and an unordered queue).  The
    per-example work is Python-bound (~0.6 ms), so `workers` > 0 assembles batches in forked
    loader PROCESSES (the memory-mapped HDF5 file is shared read-only; the large arrays come back
    through anonymous shared mappings) and still delivers them in the seeded order.
    rank / world: the ids are sharded ids[rank::world] before batching.
    copy=False (workers > 0): the large arrays of a yielded batch are windows of a shared-memory slot that
    stays untouched until TWO more batches have been requested - for consumers that stage every batch at
    once (the trainer copies it into pinned memory on receipt); saves a 15 MB copy per batch in the parent."""
    r = np.random.RandomState(seed)
    ids = list(dataset.ids)[rank::world]      # data parallelism: every rank owns a disjoint shard of the ids
    if len(ids) < batch_size:
        # (a generator: the error surfaces at the first next(), before any step is attempted)
        raise ValueError('dataset split has %d examples, fewer than batch_size=%d: no full batch can '
                         'be formed' % (len(ids), batch_size))

    def id_lists():
        e = 0
        while epochs is None or e < epochs:
            order = r.permutation(len(ids)) if shuffle else np.arange(len(ids))
            for s in range(0, len(ids) - batch_size + 1, batch_size):
                yield [ids[i] for i in order[s:s + batch_size]]
            e += 1

    if workers <= 0:
        for lst in id_lists():
            yield collate(dataset, lst)
        return
    import collections
    import mmap
    import multiprocessing as mp
    global _WORKER_DATASET, _WORKER_SLOTS
    probe = collate(dataset, ids[:1])
    per_example = sum((probe[k].nbytes + 63) // 64 * 64 for k in _BIG_KEYS)
    nslots = workers + max(lookahead, 2) + (0 if copy else 2)
    maps = [mmap.mmap(-1, per_example * batch_size + 4096) for _ in range(nslots)]
    _WORKER_DATASET = dataset
    _WORKER_SLOTS = [np.frombuffer(m, np.uint8) for m in maps]
    free = collections.deque(range(nslots))
    pending = collections.deque()
    held = collections.deque()     # copy=False: slots of the last two yielded batches

    def finish(res):
        try:
            slot, meta, b = res.get(timeout=_LOADER_TIMEOUT_S)
        except mp.TimeoutError:
            raise RuntimeError('a loader process did not deliver its batch within %d s' % _LOADER_TIMEOUT_S)
        buf = _WORKER_SLOTS[slot]
        for key, (off, shape, dt) in meta.items():
            n = int(np.prod(shape)) * np.dtype(dt).itemsize
            a = buf[off:off + n].view(dt).reshape(shape)
            b[key] = a.copy() if copy else a
        if copy:
            free.append(slot)
        else:
            held.append(slot)
            while len(held) > 2:
                free.append(held.popleft())
        return b

    with mp.get_context('fork').Pool(workers) as pool:
        for lst in id_lists():
            while not free:       # (copy=False keeps the last two yielded slots out of circulation)
                yield finish(pending.popleft())
            pending.append(pool.apply_async(_collate_job, ((free.popleft(), lst),)))
        while pending:
            yield finish(pending.popleft())
    _WORKER_SLOTS = None
