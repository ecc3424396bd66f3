#!/bin/bash
# Launch lists + in-pipeline op tables only (about 3 GPU-minutes); tools/collect_profiles.sh adds the ncu --set full pass
# and the role-stall counters.
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_infer_d4.csv python tools/profile_infer.py 16 4 > $O/pi.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train_d4.csv python tools/profile_train.py 32 4 > $O/pt.log 2>&1
python tools/time_ops.py infer 16 > $O/ops_infer.txt 2>&1
python tools/time_ops.py train 32 > $O/ops_train.txt 2>&1
python tools/bench_elementwise.py 32 > $O/elementwise.txt 2>&1
python tools/bench_attn.py > $O/attn.txt 2>&1
