#!/bin/bash
# timing + accuracy of the attention kernels with a fraction of the exp2 evaluated by the FMA-pipe polynomial
# (tuning builds painter_b200/libpk_poly_<mask>.so: -DPK_EXP_POLY_MASK=<mask> on attention_fwd.cu / attention_bwd.cu)
for lib in painter_b200/libpainter_b200.so painter_b200/libpk_poly_*.so; do
  echo "== $lib"
  PK_LIB=$PWD/$lib python scripts/time_attn_parts.py
  for c in 6 2; do PK_LIB=$PWD/$lib python scripts/gpu_check_attn.py --case $c | grep RESULT | cut -c1-420; done
done
