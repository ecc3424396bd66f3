#!/bin/bash
# Collects the round's profiling evidence on a B200 box into gpurun_out/ (copied to profiles/ afterwards).
#   launch lists (ncu gpu__time_duration, cold-cache serialised), one `ncu --set full` pass over the encoder kernels,
#   in-pipeline per-op times (CUDA events), GEMM role-stall counters (diagnostic build).
set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_infer_d4.csv python tools/profile_infer.py 16 4 > $O/pi.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train_d4.csv python tools/profile_train.py 32 4 > $O/pt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attn_fwd|layernorm" --launch-skip 30 --launch-count 16 -o $O/infer_full python tools/profile_infer.py 16 2 3 > $O/pf.log 2>&1
python tools/time_ops.py infer 16 > $O/ops_infer.txt 2>&1
python tools/time_ops.py train 32 > $O/ops_train.txt 2>&1
python tools/gemm_roles.py > $O/gemm_roles.txt 2>&1
python tools/conv_roles.py 32 > $O/conv_roles.txt 2>&1
python tools/bench_decoder_gemms.py 32 5 > $O/decoder_gemms.txt 2>&1
tail -3 $O/pi.log $O/pt.log $O/pf.log
