#!/bin/bash
for S in 64 128; do for L in 40 64; do
  echo "SEG=$S L=$L $(KJX_SEG=$S KJX_CHUNK=$L python bench.py --log2n ${1:-21} --steps 200 --warmup 5 --no-cpu --no-e2e --no-parity --no-extra | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"]*1e3,1), {k[:6]: round(v*1e3,1) for k,v in d["roofline"]["phase_ms"].items()})')"
done; done
