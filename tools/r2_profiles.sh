#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): ncu captures + bench lines, written to gpurun_out/ and copied to profiles/ afterwards.
set -x
O=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/r2_launches_step_samh.csv python tools/prof_step.py > /dev/null 2>&1
CVB_ARCH=ViT256 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/r2_launches_step_vit256.csv python tools/prof_step.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:window_tc_kernel -s 5 -c 1 -o $O/r2_full_window_tc python tools/prof_step.py > /dev/null 2>&1
CVB_ARCH=ViT256 timeout 400 ncu --set full --clock-control none --import-source on -k regex:flash_tc_kernel -s 3 -c 1 -o $O/r2_full_flash_tc_vit python tools/prof_step.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:flash_tc_kernel -s 1 -c 1 -o $O/r2_full_flash_tc_sam python tools/prof_step.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"sobel_kernel|table_accum_kernel|ccl_merge_kernel" -c 3 -o $O/r2_full_postproc python tools/prof_post.py 1 > /dev/null 2>&1
(timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1) > $O/r2_bench_c3.json
(timeout 900 python bench.py --arch ViT256 --steps 20 --warmup 3 2>&1 | tail -1) > $O/r2_bench_c2.json
(timeout 900 python bench.py --steps 400 --warmup 5 --no-cpu 2>&1 | tail -1) > $O/r2_bench_c3_steady400.json
(timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1) > $O/r2_bench_reference_arm.json
ls -la $O | tail -15
