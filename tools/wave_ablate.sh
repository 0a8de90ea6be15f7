#!/bin/bash
# needs a library built with DMG_NVCC_EXTRA=-DDMG_WAVE_ABLATION (the ablation instantiations are not in the product build)
# ablation timings of the wave scorer (profiling only): per-kernel warm durations under ncu for each DMG_WAVE_DBG mask
for d in 0 1 2 4 8 16 3 7 31; do
  DMG_WAVE_DBG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:wave_score -s 16 -c 8 --csv --log-file gpurun_out/abl_$d.csv python bench.py --steps 2 --warmup 1 --inflight 1 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/abl_$d.csv')) if len(r)>10 and r[0].isdigit()]
v=sorted(int(r[-1]) for r in rows)
print('dbg', $d, 'median score kernel ns', v[len(v)//2] if v else None, v)
PY
done
