"""Generates tests/golden/dataset_karel_golden.json by running the REFERENCE's own loader class,
karel_env/dataset_karel.py `Dataset.get_data` (and `create_default_splits` for the id split /
shuffle), on a dataset directory written by demo2program_b200.dataset.write_karel_dataset with fixed
seeds.  The reference module is imported from /root/reference (build container only) with two
stand-ins for packages that are not installed: `h5py` (File -> demo2program_b200.hdf5_lite.File,
datasets get the `.value` property the reference uses) and `colorlog` (logging cosmetics).
Stored per example: shape, dtype and sha256 of each element of the 13-tuple, plus the split ids in
the reference's shuffled order.  tests/test_hdf5.py rebuilds the same directory and checks
demo2program_b200.dataset.H5Dataset / create_default_splits against it."""
import hashlib
import json
import logging
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
SPEC = dict(n_train=10, n_test=4, n_val=3, k=6, test_k=3, seed=77)
NUM_K = 4


def digest(a):
    a = np.ascontiguousarray(a)
    return {'shape': list(a.shape), 'dtype': str(a.dtype), 'sha256': hashlib.sha256(a.tobytes()).hexdigest()}


def main():
    from demo2program_b200 import dataset as ds, hdf5_lite
    sys.dont_write_bytecode = True
    hdf5_lite.Dataset.value = property(lambda self: self[()])
    h5 = types.ModuleType('h5py')
    h5.File = lambda path, mode='r': hdf5_lite.File(path)
    sys.modules['h5py'] = h5
    cl = types.ModuleType('colorlog')
    cl.ColoredFormatter = lambda *a, **k: logging.Formatter('%(message)s')
    sys.modules['colorlog'] = cl
    sys.path.insert(0, '/root/reference')
    logging.addLevelName(15, 'INFOV')
    from karel_env import dataset_karel as ref
    with tempfile.TemporaryDirectory() as d:
        ds.write_karel_dataset(d, SPEC['n_train'], SPEC['n_test'], SPEC['n_val'], SPEC['k'],
                               test_k=SPEC['test_k'], seed=SPEC['seed'])
        tr, te, va = ref.create_default_splits(d, num_k=NUM_K)
        out = {'spec': SPEC, 'num_k': NUM_K,
               'splits': {'train': list(tr.ids), 'test': list(te.ids), 'val': list(va.ids)},
               'attrs': {'max_demo_len': tr.max_demo_len, 'max_program_len': tr.max_program_len,
                         'num_program_tokens': tr.num_program_tokens, 'num_action_tokens': tr.num_action_tokens},
               'examples': {}}
        for split in (tr, te, va):
            for i in split.ids:
                out['examples'][i] = [digest(np.asarray(x)) for x in split.get_data(i)]
    json.dump(out, open(os.path.join(HERE, 'dataset_karel_golden.json'), 'w'), separators=(',', ':'))
    print('examples', len(out['examples']), 'tuple length', len(next(iter(out['examples'].values()))))
    print(out['splits']['train'][:3])


if __name__ == '__main__':
    main()
