#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (K2 subset, all geometries incl. dense)"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "synth_parity or edge_shapes or mixed or gate_mix or constant_wire or full_size" 2>&1 | tail -4
B="--steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 --no-qap --no-one-shot --no-overlap"
show() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']
        print('isolated %.5f ms frac %.4f  stream MB %.1f' % (r['kernel_ms_mean'], r['frac'], r['device_stream_bytes_per_launch']/1e6))
    elif 'rror' in l: print(l.strip())
"; }
echo "=== dense (auto geometry)"; timeout 600 python bench.py $B --dense 2>&1 | show
echo "=== dense variant 6"; timeout 600 python bench.py $B --dense --variant 6 2>&1 | show
echo "=== default sparse with variant 8"; timeout 600 python bench.py $B --variant 8 2>&1 | show
echo "=== default sparse"; timeout 600 python bench.py $B 2>&1 | show
