"""Generates tests/golden/texture_hdf5.json from the reference's only HDF5 file
(/root/reference/karel_env/asset/texture.hdf5, written by h5py): dataset names, shapes, dtypes,
content hashes and simple statistics as read by demo2program_b200/hdf5_lite.py.  The values were
cross-checked by hand against the raw file (contiguous float64 little-endian payloads at the
addresses named by the object headers); run here, where /root/reference is mounted."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from demo2program_b200 import hdf5_lite

out = {}
with hdf5_lite.File('/root/reference/karel_env/asset/texture.hdf5') as f:
    for name in sorted(f.keys()):
        a = np.ascontiguousarray(f[name][()])
        out[name] = {'shape': list(a.shape), 'dtype': str(a.dtype), 'sha256': hashlib.sha256(a.tobytes()).hexdigest(),
                     'sum': float(a.sum()), 'min': float(a.min()), 'max': float(a.max())}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'texture_hdf5.json'), 'w'), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
