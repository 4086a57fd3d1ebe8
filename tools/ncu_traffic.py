"""profiles/ncu_traffic.json from a committed `ncu --set full` summary (the text written next to the
.ncu-rep by the profile scripts: blocks of `Kernel Name ...`, `dram__bytes_read.sum X Mbyte`,
`dram__bytes_write.sum Y Mbyte`).  bench.py reads the JSON for `roofline.traffic` (bytes per launch of the
dominant kernel); the first capture of each kernel in the file is used.

usage: python tools/ncu_traffic.py profiles/r03i_ncu_full_lstm_persist_fwd.txt [more summaries ...]"""
import json
import os
import re
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def parse(path):
    out, name, rd = {}, None, None
    for line in open(path):
        m = re.match(r'Kernel Name\s+(.*\S)', line)
        if m:
            name, rd = m.group(1), None
            continue
        m = re.match(r'dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)', line)
        if m and name:
            v = float(m.group(2)) * UNIT[m.group(3)]
            if m.group(1) == 'read':
                rd = v
            elif rd is not None and name not in out:
                out[name] = {'dram_bytes_per_launch': rd + v, 'dram_read': rd, 'dram_write': v,
                             'source': os.path.relpath(path, os.path.join(os.path.dirname(__file__), '..'))}
    return out


if __name__ == '__main__':
    tab = {}
    for p in sys.argv[1:]:
        for k, v in parse(p).items():
            tab.setdefault(k, v)
    dst = os.path.join(os.path.dirname(__file__), '..', 'profiles', 'ncu_traffic.json')
    with open(dst, 'w') as f:
        json.dump(tab, f, indent=1, sort_keys=True)
    print(json.dumps(tab, indent=1, sort_keys=True))
