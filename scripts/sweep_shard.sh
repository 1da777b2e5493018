#!/bin/bash
# knob sweep of the linked iteration on one 2^21-step shard (the per-GPU size of the 8-GPU run)
for T in 32 64 128; do for L in 28 32 36 40 48 64; do
  echo "T=$T L=$L $(KJX_THREADS=$T KJX_CHUNK=$L python bench.py --log2n ${1:-21} --steps 200 --warmup 5 --no-cpu --no-e2e --no-parity | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"]*1e3,1), {k[:6]: round(v*1e3,1) for k,v in d["roofline"]["phase_ms"].items()})')"
done; done
