#!/bin/bash
# FFN kernel timing under the TW_FFN_DBG experiments (bring-up tool)
for d in 0 1 2 4 3 5 6 7; do
  echo -n "TW_FFN_DBG=$d  "
  TW_FFN_DBG=$d python bench.py --steps 2 --warmup 3 --chains 1024 --precision bf16x3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ffn avg ms', round(d['roofline']['avg_launch_ms'],4), 'step ms', round(d['ms_per_step'],2))"
done
