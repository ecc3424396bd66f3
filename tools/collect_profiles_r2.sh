#!/bin/bash
# Round-2 profiling evidence on a B200 box, written to gpurun_out/ (copied to profiles/r02_* afterwards):
#   1. ncu launch list of the bench command itself (gpu__time_duration, cold-cache, serialised)
#   2. ncu launch lists of one eval forward / one training step at depth 4 (every kernel type once per block)
#   3. `ncu --set full` over the eval-forward kernels and over the training-only kernels (attention backward, decoder,
#      loss / optimiser / re-layout), summarised with tools/ncu_summary.py
#   4. in-pipeline per-op times (CUDA events around every ops.* call, all 40 blocks), attention / elementwise microbenches
O=gpurun_out
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_bf16|attn_|layernorm|upsample|tokens_to_map|prep_|patch_|fill_prefix|heads_|bn_|gram32" -c 1200 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > $O/r02_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_infer_d4.csv python tools/profile_infer.py 16 4 > $O/pi.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_train_d4.csv python tools/profile_train.py 32 4 > $O/pt.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"gemm_bf16|attn_fwd|layernorm|upsample|tokens_to_map|prep_|patch_" --launch-skip 40 --launch-count 36 -o $O/r02_infer_full python tools/profile_infer.py 16 2 3 > $O/pf.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"attn_bwd|layernorm_bwd|bn_relu|heads_|gram32|loss_|adam|sumsq|norm_fin|gather_cast|lora_|transpose|f16_to|zero_insert|add_bf16|upsample2x_bwd|tokens_to_map_bwd" --launch-skip 60 --launch-count 56 -o $O/r02_train_full python tools/profile_train.py 32 2 > $O/ptf.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"gemm_bf16" --launch-skip 150 --launch-count 40 -o $O/r02_train_gemm_full python tools/profile_train.py 32 2 > $O/ptg.log 2>&1
python tools/ncu_summary.py $O/r02_infer_full.ncu-rep $O/r02_ncu_full_infer_b16.csv > /dev/null 2>&1
python tools/ncu_summary.py $O/r02_train_full.ncu-rep $O/r02_ncu_full_train_b32.csv > /dev/null 2>&1
python tools/ncu_summary.py $O/r02_train_gemm_full.ncu-rep $O/r02_ncu_full_train_gemms_b32.csv > /dev/null 2>&1
rm -f $O/*.ncu-rep   # the reports exceed what gpurun copies back; the per-kernel CSV summaries above are what is kept
python tools/summarize_launches.py $O/r02_launches_infer_d4.csv > $O/r02_launches_infer_d4.txt 2>&1
python tools/summarize_launches.py $O/r02_launches_train_d4.csv lora_refresh > $O/r02_launches_train_d4.txt 2>&1
python tools/summarize_launches.py $O/r02_launches_bench.csv > $O/r02_launches_bench.txt 2>&1
timeout 600 python tools/time_ops.py infer 16 > $O/r02_ops_infer.txt 2>&1
timeout 600 python tools/time_ops.py train 32 > $O/r02_ops_train.txt 2>&1
timeout 300 python tools/bench_elementwise.py 32 > $O/r02_elementwise.txt 2>&1
timeout 300 python tools/bench_attn.py > $O/r02_attn.txt 2>&1
ls -la $O/r02_*
for f in $O/pi.log $O/pt.log $O/pf.log $O/ptf.log $O/ptg.log $O/r02_bench_under_ncu.log; do tail -n 2 $f; done
du -sh $O
