"""Developer tool (CPU): throughput of the HDF5 dataset path (hdf5_lite reader + the restated
get_data / collate of karel_env/dataset_karel.py) on a synthetic dataset directory."""
import sys
import tempfile
import time

sys.path.insert(0, '.')
from demo2program_b200 import dataset as ds

n, k, B = 512, 10, 32
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    ds.write_karel_dataset(d, n, 32, 32, k, seed=0)
    t1 = time.perf_counter()
    tr, _, _ = ds.create_default_splits(d, num_k=k)
    t2 = time.perf_counter()
    print('write %d examples: %.2f s; open + split: %.3f s' % (n + 64, t1 - t0, t2 - t1))
    for workers, copy in ((0, True), (4, True), (4, False), (7, False)):
        t2 = time.perf_counter()
        nb = 0
        for b in ds.batches(tr, B, shuffle=True, seed=0, epochs=12, workers=workers, copy=copy):
            nb += 1
        t3 = time.perf_counter()
        print('workers=%d copy=%s: %d batches of %d (k=%d): %.1f examples/s, %.2f ms/batch'
              % (workers, copy, nb, B, k, nb * B / (t3 - t2), 1e3 * (t3 - t2) / nb))
