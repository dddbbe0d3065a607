#!/bin/bash
# round 2 evidence: launch lists, --set full captures, fusion roofline, latency sweep, fit, adapnet
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/r02_bench.err; tail -c 200 $O/r02_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $O/r02_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --soak 0 --no-extras > $O/r02_bench_under_ncu.log 2>&1
# --set full captures: exported to CSV on the box (the .ncu-rep files exceed what gpurun copies back)
timeout 900 ncu --set full --clock-control none -k 'regex:conv_igemm|conv_c1|head_fused|decode' -s 32 -c 16 -o /tmp/r02_stream python tools/perf_probe.py 16 1 3 > $O/r02_ncu_stream.log 2>&1
ncu -i /tmp/r02_stream.ncu-rep --page raw --csv > $O/r02_stream_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k 'regex:decode_bayes' -s 3 -c 1 -o /tmp/r02_tail python bench.py --steps 1 --warmup 3 --soak 0 --no-extras > $O/r02_ncu_tail.log 2>&1
ncu -i /tmp/r02_tail.ncu-rep --page raw --csv > $O/r02_tail_raw.csv 2>/dev/null
timeout 600 python tools/fusion_bench.py > $O/r02_fusion_roofline.txt 2>&1; cp $O/fusion_roofline.json $O/r02_fusion_roofline.json
timeout 900 python tools/timing.py --repetitions 30 --graph --cpu --json $O/r02_timing_sweep.json > $O/r02_timing_sweep.txt 2>&1; tail -12 $O/r02_timing_sweep.txt
timeout 600 python tools/fit_bench.py --steps 10 2>&1 | tail -1 > $O/r02_fit_bench.json; cat $O/r02_fit_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_fit_launches.csv python tools/fit_bench.py --steps 1 --warmup 1 > $O/r02_fit_under_ncu.log 2>&1
timeout 300 python tools/adapnet_bench.py 16 10 2>&1 | tail -1 > $O/r02_adapnet_bench.json; cat $O/r02_adapnet_bench.json
ls -la $O | grep r02_ | tail -20
