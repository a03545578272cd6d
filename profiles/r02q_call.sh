#!/bin/bash
# single GPU, A/B only: k_fwd with the L2 prefetch of the replacing CTA's particle lines (FWD_AHEAD), k_grid_op CTA counts and the warp-per-block variant,
# backward kernels with the same prefetch-ahead, cell-sort period
set -x
tag=${1:-r02q}
mkdir -p gpurun_out
for v in a_default b_ahead1184 c_ahead592 d_ahead2368 g_gop4 h_gop16 i_gopwarp; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/${tag}_ab.jsonl
  FMPM_LIB="$PWD/gpurun_variants/$v.so" AB_STEPS=40 timeout 200 python profiles/fwd_ab.py 3 >> gpurun_out/${tag}_ab.jsonl 2>> gpurun_out/${tag}_ab.err
done
for srt in 2 8; do
  echo "{\"variant\": \"a_default sort_every=$srt\"}" >> gpurun_out/${tag}_ab.jsonl
  FMPM_LIB="$PWD/gpurun_variants/a_default.so" AB_SORT=$srt AB_STEPS=40 timeout 200 python profiles/fwd_ab.py 3 >> gpurun_out/${tag}_ab.jsonl 2>> gpurun_out/${tag}_ab.err
done
cat gpurun_out/${tag}_ab.jsonl | cut -c1-120
for v in a_default e_pg592 f_pg592_sc888; do
  FMPM_LIB="$PWD/gpurun_variants/$v.so" timeout 200 python profiles/bwd_ab.py >> gpurun_out/${tag}_bwd_ab.jsonl 2>> gpurun_out/${tag}_ab.err
done
cat gpurun_out/${tag}_bwd_ab.jsonl
tail -3 gpurun_out/${tag}_ab.err
