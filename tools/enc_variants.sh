#!/bin/bash
# Encoder layout / block-size experiments: bench.py --no-extras per variant (value, e2e, parity), one line each.
mkdir -p gpurun_out
for lib in compression_b200/libtfcb200.so compression_b200/libtfcb200_b128.so; do
  [ -f "$lib" ] || continue
  for rot in 0 1 2 6; do
    TFCB_LIB_PATH=$PWD/$lib TFCB_ENC_ROT=$rot python bench.py --no-extras --steps 100 2>/dev/null | \
      python -c "import json,sys; d=json.load(sys.stdin); print('$lib rot=$rot value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parity', d['parity_checked'])"
  done
done
