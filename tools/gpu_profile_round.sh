# Round profiling pass (run on a GPU box from the repo root): plain runs first, then the same commands under ncu.
set -x
python -m pytest tests/test_gpu_parity.py -q -s -k "body" --timeout 600 2>&1 | grep -E "^\[|passed|failed|^E " | head -20
python bench.py --steps 20 --warmup 5 --no-saturated > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-saturated --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
python tools/gpu_prof_step.py latency 4096 24 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 20 -c 1 -f -o gpurun_out/r2_latency python tools/gpu_prof_step.py latency 4096 24 > gpurun_out/r2_ncu1.log 2>&1
python tools/gpu_prof_step.py throughput 65536 24 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 20 -c 1 -f -o gpurun_out/r2_throughput python tools/gpu_prof_step.py throughput 65536 24 > gpurun_out/r2_ncu2.log 2>&1
python tools/gpu_prof_step.py latency 4096 124 1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 120 -c 1 -f -o gpurun_out/r2_body python tools/gpu_prof_step.py latency 4096 124 1 > gpurun_out/r2_ncu3.log 2>&1
python tools/gpu_prof_gae.py && ncu --set full --clock-control none --import-source on -k regex:gae -s 2 -c 1 -f -o gpurun_out/r2_gae python tools/gpu_prof_gae.py > gpurun_out/r2_ncu4.log 2>&1
for r in r2_latency r2_throughput r2_body r2_gae; do python tools/ncu_raw_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.summary.txt 2>&1; done
rm -f gpurun_out/r2_body.ncu-rep gpurun_out/r2_gae.ncu-rep   # the merge back is limited to 64 MiB: keep the two step-kernel reports
ls -la gpurun_out/*.ncu-rep
python tools/gpu_ab.py 16384,65536 solorl_b200/libsolo_b200.so tools/_ab/libsolo_tp_6_2.so tools/_ab/libsolo_tp_5_2.so tools/_ab/libsolo_tp_4_3.so tools/_ab/libsolo_tp_12_1.so 2>&1 | grep solo12 | tee gpurun_out/r2_tp_shapes.txt
